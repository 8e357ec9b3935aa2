import sys, json, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import _cases as CS
import mcpower_b200 as M
G = CS.crossing_goldens()
for name, g in G.items():
    s = CS.crossing_spec(g["case"])
    t=time.time()
    r = M.find_sample_size(s, g["from_size"], g["to_size"], g["n_sims"], M.splitmix64(g["seed"]), by=g["by"])
    dt=time.time()-t
    got = [(f["status"], f.get("n_achievable")) for f in r["fitted"]]
    print(name, "golden", [(d["status"], d["n_achievable"]) for d in g["dense"].values()], "got", got, "atom", r["cluster_atom"], "pts", len(r["grid_or_trace"]), f"{dt:.1f}s kernel {r['kernel_ms']:.0f} ms")
