from hpcs import unpatched

hyp_lca = unpatched("hpcs.distances.lca.hyp_lca")
