from .hyp_embed import ExpMap, MLPExpMap
