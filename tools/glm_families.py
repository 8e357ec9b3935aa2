import sys; sys.path.insert(0, '/root/repo')
import mcpower_b200 as M
from mcpower_b200 import spec as S
seed = M.splitmix64(2137)
for fam in ("logit", "probit", "poisson"):
    s = S.logit_spec()
    if fam == "probit":
        s["link"] = "probit"; s["effect_sizes"][0] = -0.5
    elif fam == "poisson":
        s["outcome_kind"] = "count"; s["effect_sizes"] = [0.7, 0.3, 0.2, 0.1, 0.1]
    grid = [50, 100, 200, 500, 1000, 2000]
    M.run_batch(s, grid, 4096, seed)
    r = M.run_batch(s, grid, 1 << 17, seed)
    print(fam, "%.3e fits/s" % ((1 << 17) * 6 / (r.kernel_ms * 1e-3)), "conv %.4f" % (r.counters[:, 2 * 4].sum() / (6 * (1 << 17))))
