from hpcs import unpatched

get_optimal_k = unpatched("hpcs.utils.scores.get_optimal_k")
