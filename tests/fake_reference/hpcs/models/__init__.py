from .shapenet_hyp_hc import ShapeNetHypHC
from .partnet_hyp_hc import PartNetHypHC
