from .cosine import CosineSimilarity
from .lca import hyp_lca
