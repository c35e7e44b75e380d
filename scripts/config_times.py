"""Step times of the BASELINE.json configs on ONE GPU (the bench line is C2; these are the other
configs, timed the same way: device-resident model, CUDA events around `steps` epochs).

  python scripts/config_times.py [C1 C2 C3 C4a C4b C5] [--steps 10]

C5 is the per-rank shard of the 8-GPU case (10 000 of the 80 000 samples x 50 000 features, K=128)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.simulate import C2_BLOCKS, scale_blocks, simulate_problem


def graphs(N, K, n_edges, n_virtual, rng):
    out = []
    for k in range(K):
        a = rng.integers(1, N + 1, size=n_edges); b = rng.integers(1, N + 1, size=n_edges)
        s = rng.choice([-1.0, 1.0], size=n_edges)
        el = [[int(x), int(y), float(z)] for x, y, z in zip(a, b, s) if x != y]
        for v in range(n_virtual):
            el.append([int(rng.integers(1, N + 1)), f"virt{k}_{v}", 1.0])
        out.append(el)
    return out


def feature_sets(blocks, n_sets, rng):
    sets, c0 = {}, 0
    for v, d, n in blocks:
        sets[v] = [sorted(int(c0 + 1 + j) for j in rng.choice(n, size=int(rng.integers(10, min(200, n))), replace=False))
                   for _ in range(n_sets)]
        c0 += n
    return sets


def build(name, rng):
    if name == "C1":
        return simulate_problem(100, blocks=(("mrnaseq", "normal", 200),), K=8, seed=1, missing=0.0), "100x200 K=8 normal"
    if name == "C2":
        return simulate_problem(10000, blocks=C2_BLOCKS, K=64, seed=2, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0)), \
            "10000x30000 K=64 mixed, 30% missing"
    if name == "P25":
        return simulate_problem(10000, blocks=C2_BLOCKS, K=25, seed=2, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0)), \
            "C2 shape at the reference's production latent dimension K=25 (fit_matfac.jl:150)"
    if name in ("C3", "C3s", "C3b1"):
        views = [b[0] for b in C2_BLOCKS]
        return simulate_problem(10000, blocks=C2_BLOCKS, K=64, seed=3, missing=0.3, batch_views=views, n_batches=1 if name == "C3b1" else 40,
                                n_conditions=20, sort_batches=name == "C3s"), \
            "C2 + 40 batches x 4 views of shift/scale + 20 sample conditions (group reg on X)" + \
            ("; samples listed batch by batch" if name == "C3s" else "; batch ids iid per sample and view")
    if name == "C4a":
        N, K = 30000, 256
        g = graphs(N, K, 7800, 780, rng)
        return simulate_problem(10000, blocks=(("mrnaseq", "normal", N),), K=K, seed=4, missing=0.3,
                                model_kwargs=dict(feature_graphs=g, lambda_Y_graph=1.0, lambda_Y_selective_l1=0.5)), \
            f"10000x30000 K=256, per-factor graphs: {sum(len(x) for x in g)} edges, 780 virtual nodes each, selective L1"
    if name == "C4b":
        N, K = 30000, 256
        blocks = (("methylation", "normal", 15000), ("mrnaseq", "normal", 15000))
        fs = feature_sets(blocks, 500, rng)
        return simulate_problem(10000, blocks=blocks, K=K, seed=4, missing=0.3,
                                model_kwargs=dict(Y_fsard=True, feature_sets_dict=fs)), \
            "10000x30000 K=256, feature-set ARD on Y (500 sets per view)"
    if name == "C5":
        blocks = (("mutation", "bernoulli", 20000), ("mrnaseq", "normal", 30000))
        return simulate_problem(10000, blocks=blocks, K=128, seed=5, missing=0.3, model_kwargs=dict(lambda_X_l2=1.0)), \
            "per-rank shard of 80000x50000 K=128 at 8 GPUs: 10000x50000"
    raise SystemExit(f"unknown config {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C1", "C2", "C3", "C4a", "C4b", "C5"])
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    for name in args.configs:
        t0 = time.time()
        model, desc = build(name, rng)
        t_build = time.time() - t0
        eng = P.Engine(model)
        eng.reset_opt_state(1e-8)
        common = dict(lr=0.05, update_X=1, update_Y=1, update_col_layers=1, no_terminate=1, check_every=1 << 20,
                      rel_tol=0.0, abs_tol=0.0)
        eng.fit(eng.make_opts(epoch=1, max_epochs=3, **common))
        eng.set_profiling(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = eng.torch_stream()
        torch.cuda.synchronize()
        ev0.record(stream)
        h = eng.fit(eng.make_opts(epoch=4, max_epochs=3 + args.steps, **common))
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / args.steps
        n, dp_ms, dp_min = eng.get_profile()
        M, N = model.data.shape
        K = model.matfac.X.shape[0]
        print(json.dumps({"config": name, "workload": desc, "ms_per_step": ms, "iters_per_s": 1e3 / ms,
                          "data_pass_ms": dp_ms, "data_pass_GBs": 4.0 * M * N / (dp_ms * 1e-3) / 1e9 if dp_ms else None,
                          "kernel_launches_per_step": h["kernel_launches"] / args.steps, "host_build_s": round(t_build, 1),
                          "loss_first_last": [h["loss"][0], h["loss"][-1]], "M": M, "N": N, "K": K}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
