from .ar_model import ARModel
from .base_graph_model import BaseGraphModel
from .base_hi_graph_model import BaseHiGraphModel
from .graph_lam import GraphLAM
from .hi_lam import HiLAM
from .hi_lam_parallel import HiLAMParallel

MODELS = {"graph_lam": GraphLAM, "hi_lam": HiLAM, "hi_lam_parallel": HiLAMParallel}
