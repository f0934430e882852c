"""B200-native (sm_100a) implementation of Neural-LAM's InteractionNet
message-passing hot path and the GraphLAM / HiLAM / HiLAMParallel train step
around it.  Module layout mirrors /root/reference/neural_lam/ for the path:

    interaction_net.InteractionNet   <- interaction_net.py:10-163
    utils.make_mlp / load_graph      <- utils.py:11-214
    models.{GraphLAM,HiLAM,HiLAMParallel} <- models/*.py
    metrics.{wmse,mse}               <- metrics.py:21-113
    create_graph                     <- create_graph.py (input format)
    lib                              <- ctypes binding of csrc/libnlam_b200.so

Compute goes through the C-ABI library built from csrc/ (hand-written CUDA,
no CPU fallback: ops raise if the library is missing).
"""
__version__ = "0.1.0"
