import sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as e; e.build()
from neural_lam_b200 import ops, utils
dev = torch.device('cuda:0')
ops.set_precision('bf16')
torch.manual_seed(0)
for dout in (17,):
    mlp = utils.make_mlp([64, 64, dout], layer_norm=False).to(dev)
    x = torch.randn(4, 63784, 64, device=dev, requires_grad=True)
    w = torch.randn(4, 63784, dout, device=dev)
    for _ in range(2):
        (mlp(x) * w).sum().backward()
    torch.cuda.synchronize()
    t = ops.KernelTimer(); ops.set_timer(t)
    for _ in range(3):
        (mlp(x) * w).sum().backward()
    torch.cuda.synchronize(); ops.set_timer(None)
    for tag, (n, ms, nb, fl) in sorted(t.summary().items(), key=lambda kv: -kv[1][1]):
        print(f"  {ms/n*1e3:9.1f} us  {tag}")
