"""Pin the CPU oracle against the reference's own known-answer tests
(/root/reference/test/runtests.jl, transcribed to tests/golden/runtests_known_answers.json)."""
import json
import os

import numpy as np
import pytest

from oracle import pmf_oracle as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "runtests_known_answers.json")))


def jr(pair):
    """Julia 1-based inclusive [a, b] -> Python range."""
    return range(pair[0] - 1, pair[1])


def test_is_contiguous():
    for c in G["is_contiguous"]["cases"]:
        assert O.is_contiguous(c["vec"]) == c["expect"]


def test_ids_to_ranges():
    for c in G["ids_to_ranges"]["cases"]:
        assert O.ids_to_ranges(c["vec"]) == [jr(p) for p in c["expect"]]


def test_subset_ranges():
    for c in G["subset_ranges"]["cases"]:
        new, lo, hi = O.subset_ranges([jr(p) for p in c["ranges"]], jr(c["rng"]))
        assert new == [jr(p) for p in c["expect"]["ranges"]]
        assert (lo + 1, hi + 1) == (c["expect"]["r_min"], c["expect"]["r_max"])


def test_ids_to_ind_mat():
    g = G["ids_to_ind_mat"]
    assert np.array_equal(O.ids_to_ind_mat(g["vec"]), np.array(g["expect"], dtype=bool))


def test_keymatch():
    g = G["keymatch"]
    li, ri = O.keymatch(g["l_keys"], g["r_keys"])
    assert [i + 1 for i in li] == g["l_idx"] and [i + 1 for i in ri] == g["r_idx"]


def test_nanstats():
    g = G["nanstats"]
    v = [np.nan if x is None else x for x in g["vec"]]
    assert O.nansum(v) == g["nansum"] and O.nanmean(v) == g["nanmean"] and O.nanvar(v) == g["nanvar"]


def test_edgelist_to_spmat():
    g = G["edgelist_to_spmat"]
    n2i = {c: i for i, c in enumerate(g["nodes"])}
    sp = O.edgelist_to_spmat(g["edgelist"], n2i, epsilon=g["epsilon"])
    assert np.allclose(sp.toarray(), np.array(g["expect"]))


def _ba(g, key_cols="col_batches"):
    vals = []
    ranges = O.ids_to_ranges(g[key_cols])
    for vd, cr in zip(g["values"], ranges):
        vals.append({int(k): np.full(len(cr), v) for k, v in vd.items()})
    return O.BatchArray.construct(g[key_cols], g["row_batches"], vals)


def test_batch_array_ctor_view_zero():
    g = G["batch_array"]
    ba = _ba(g)
    assert ba.col_ranges == [jr(p) for p in g["col_ranges"]]
    for rb, e in zip(ba.row_batches, g["indicators"]):
        assert np.array_equal(rb, np.array(e, dtype=bool))
    for v, e in zip(ba.values, g["values_expect"]):
        assert np.array_equal(v, np.array(e))
    bv = ba.view(jr(g["view_rows"]), jr(g["view_cols"]))
    assert bv.col_ranges == [jr(p) for p in g["view_col_ranges"]]
    for rb, e in zip(bv.row_batches, g["indicators"][:2]):
        assert np.array_equal(rb, np.array(e, dtype=bool)[1:4])
    for v, e in zip(bv.values, g["view_values"]):
        assert np.array_equal(v, np.array(e))
    # Vector{Int} rows and nested views
    bv2 = ba.view(np.array([1, 2, 3]), jr(g["view_cols"]))
    assert bv2.col_ranges == bv.col_ranges
    bvv = bv.view(range(0, 2), range(0, 3))
    assert bvv.col_ranges[0] == range(0, 2)
    # gap columns / empty view
    gg = dict(g, values=g["values"])
    gappy = _ba(gg, "gappy_col_batches")
    assert gappy.col_ranges == [jr(p) for p in g["gappy_col_ranges"]]
    ev = gappy.view(range(0, 5), jr(g["gappy_empty_view_cols"]))
    assert ev.col_ranges == [] and ev.values == [] and ev.row_batches == []
    z = ba.zero()
    assert all(np.all(v == 0) for v in z.values) and z.col_ranges == ba.col_ranges


def test_batch_array_arithmetic_and_pullbacks():
    g = G["batch_array"]
    ba = _ba(g)
    A = np.zeros((5, 7))
    test_mat = np.array(g["add_expect"])
    assert np.array_equal(ba.add_to(A), test_mat)
    A_bar, vb = ba.add_pullback(np.ones((5, 7)))
    assert np.array_equal(A_bar, np.ones((5, 7)))
    for b, e in zip(vb, g["grad_sum_values"]):
        assert np.array_equal(b, np.array(e, dtype=float))
    other = test_mat.copy()
    other[:, 3] = 1
    ones = np.ones((5, 7))
    assert np.allclose(ba.mul_to(ones), other)
    A_bar, vb = ba.mul_pullback(ones, np.ones((5, 7)))
    assert np.allclose(A_bar, other)
    for b, e in zip(vb, g["grad_sum_values"]):
        assert np.array_equal(b, np.array(e, dtype=float))
    # exp and its gradient n_b * exp(v)  (runtests.jl:229-236)
    assert np.array_equal(ba.exp().mul_to(ones), np.exp(test_mat))
    eb = ba.exp()
    _, vb = eb.mul_pullback(ones, np.ones((5, 7)))
    for b, ev, e, v in zip(vb, eb.values, g["grad_sum_values"], ba.values):
        assert np.allclose(b * ev, np.array(e, dtype=float) * np.exp(v))
    res = O.ba_map(lambda a: a, ba, test_mat)
    for r, e in zip(res, g["ba_map_identity"]):
        assert np.allclose(r, np.array(e))


def test_layers_forward_and_grads():
    """test/runtests.jl:348-453 (inputs there are unseeded randn; the assertions
    are formulas, re-evaluated here on seeded inputs)."""
    rng = np.random.default_rng(0)
    M, N, K = 20, 30, 4
    X, Y = rng.standard_normal((K, M)), rng.standard_normal((K, N))
    xy = X.T @ Y
    col_batches = ["colbatch1"] * 15 + ["colbatch2"] * 15
    rbs = [f"rowbatch{i}" for i in range(1, 5) for _ in range(5)]
    bd = {"colbatch1": rbs, "colbatch2": rbs}
    vds = [{b: np.zeros(15) for b in O.unique_in_order(rbs)} for _ in range(2)]
    ld = O.BatchArray.construct(col_batches, bd, vds)
    for v in ld.values:
        v[...] = rng.standard_normal(v.shape) * 0.3
    th = ld.copy()
    nm = O.NoiseModel.from_distributions(["normal"] * N)
    logsigma, mu = rng.standard_normal(N) * 0.2, rng.standard_normal(N)
    m = O.OracleModel(X=X, Y=Y, logsigma=logsigma, mu=mu, logdelta=ld, theta=th, noise=nm)
    # forward = fixed composition order
    Z = xy * np.exp(logsigma)[None, :]
    Z = Z * (ld.row_batches[0].astype(float) @ np.exp(np.hstack(ld.values)))
    Z = Z + mu[None, :] + th.row_batches[0].astype(float) @ np.hstack(th.values)
    assert np.allclose(O.forward(m), Z)
    # BatchShift grads of sum(f(x)): theta grad = batch size, Z grad = ones  (:412-414)
    A_bar, vb = th.add_pullback(np.ones((M, N)))
    assert all(np.array_equal(b, np.full((4, 15), 5.0)) for b in vb) and np.array_equal(A_bar, np.ones((M, N)))
    # BatchScale grads (:401-403)
    ed = ld.exp()
    A_bar, vb = ed.mul_pullback(xy, np.ones((M, N)))
    assert np.allclose(A_bar, ed.mul_to(np.ones((M, N))))
    assert np.allclose((vb[-1] * ed.values[-1])[-1, :],
                       (xy[15:20, 15:30] * np.exp(ld.values[-1][-1, :])[None, :]).sum(axis=0))
    # frozen layer => no gradient through the optimiser, loss unchanged
    D = Z + 0.1
    out = O.data_loss_grads(m, D)
    assert np.isclose(out["loss"], 0.5 * M * N * 0.01)
    # ColScale quirk: dlogsigma_j = sum_i sigma_j * Gbar_ij
    G1 = (Z - D) * (ld.row_batches[0].astype(float) @ np.exp(np.hstack(ld.values)))
    assert np.allclose(out["dlogsigma"], (np.exp(logsigma)[None, :] * G1).sum(axis=0))
    assert np.allclose(out["dmu"], (Z - D).sum(axis=0))


def test_data_grads_match_finite_differences():
    """Everything except the documented ColScale quirk must be the true derivative."""
    m, D, _ = O.simulate_model(12, {"mutation": ("bernoulli", 5), "methylation": ("normal", 6),
                                    "counts": ("poisson", 4)}, K=3, seed=3,
                               batch_views=["methylation", "counts"], n_batches=3, missing=0.2)
    out = O.data_loss_grads(m, D)
    eps = 1e-6

    def fd(arr, idx):
        old = arr[idx]
        arr[idx] = old + eps
        lp = O.data_loss_grads(m, D, want_grads=False)["loss"]
        arr[idx] = old - eps
        lm = O.data_loss_grads(m, D, want_grads=False)["loss"]
        arr[idx] = old
        return (lp - lm) / (2 * eps)

    for idx in [(0, 0), (2, 7), (1, 11)]:
        assert np.isclose(out["dX"][idx], fd(m.X, idx), rtol=1e-5, atol=1e-6)
    for idx in [(0, 0), (2, 7), (1, 14)]:
        assert np.isclose(out["dY"][idx], fd(m.Y, idx), rtol=1e-5, atol=1e-6)
    for j in [0, 6, 14]:
        assert np.isclose(out["dmu"][j], fd(m.mu, j), rtol=1e-5, atol=1e-6)
    for v in range(2):
        assert np.isclose(out["dtheta"][v][1, 2], fd(m.theta.values[v], (1, 2)), rtol=1e-5, atol=1e-6)
        assert np.isclose(out["dlogdelta"][v][1, 1], fd(m.logdelta.values[v], (1, 1)), rtol=1e-5, atol=1e-6)


def test_row_minibatching_is_exact():
    m, D, _ = O.simulate_model(40, {"methylation": ("normal", 9), "counts": ("poisson", 4)}, K=3, seed=4,
                               batch_views=["methylation"], n_batches=4, missing=0.3)
    a = O.data_loss_grads(m, D, capacity=10 ** 8)
    b = O.data_loss_grads(m, D, capacity=13 * 7)
    assert np.isclose(a["loss"], b["loss"])
    for k in ("dX", "dY", "dmu", "dlogsigma"):
        assert np.allclose(a[k], b[k])
    assert np.allclose(a["dtheta"][0], b["dtheta"][0])


def test_network_regularizer_blocks_and_values():
    g = G["network_regularizer"]
    p = g["path"]
    nr = O.NetworkRegularizer(p["data_features"], p["edgelists"])
    assert len(nr.AA) == 2
    assert np.allclose(nr.AA[0].toarray(), p["AA1"]) and np.allclose(nr.AB[0].toarray(), p["AB1"])
    assert np.allclose(nr.BB[0].toarray(), p["BB1"])
    nr2 = O.NetworkRegularizer(p["all_observed_features"], p["edgelists"])
    assert nr2.AA[0].shape == (4, 4) and nr2.AB[0].shape == (4, 0) and nr2.BB[0].shape == (0, 0)
    s = g["star"]
    ns = O.NetworkRegularizer(s["data_features"], s["edgelists"])
    assert np.allclose(ns.AA[0].toarray(), s["AA1"]) and np.allclose(ns.AB[0].toarray(), s["AB1"])
    assert np.allclose(ns.BB[0].toarray(), s["BB1"])
    d = g["derived"]
    loss, grad = ns.value_grad(np.array([d["star_y"]]))
    assert grad.shape == (1, 3)
    assert np.isclose(loss, d["star_loss"], atol=1e-6) and np.allclose(grad[0], d["star_grad"], atol=1e-6)
    assert np.isclose(-ns.x_virtual[0][0], -d["star_u"], atol=1e-6) or np.isclose(ns.x_virtual[0][0], d["star_u"], atol=1e-6)
    np1 = O.NetworkRegularizer(p["data_features"], p["edgelists"][:1])
    loss, grad = np1.value_grad(np.array([d["path_y"]]))
    assert np.isclose(loss, d["path_loss"], atol=1e-6) and np.allclose(grad[0], d["path_grad"], atol=1e-6)
    assert np.isclose(np1.value(np.array([d["path_y"]])), d["path_loss"], atol=1e-6)
    # old suite, epsilon = 0
    n0 = O.NetworkRegularizer(s["data_features"], s["edgelists"], epsilon=0.0)
    loss, grad = n0.value_grad(np.array([[1.0, 1.0, 1.0]]))
    # with eps=0: AA=I, AB=-1, BB=3: u = 1, loss = 0.5*3 - 3 + 1.5 = 0 ; the old suite used
    # y=[1,0,0]-style inputs; keep only the structural check here
    assert np.isfinite(loss)


def test_selective_l1():
    g = G["selective_l1"]
    reg = O.SelectiveL1Reg(g["data_features"], g["edgelists"])
    assert np.array_equal(reg.l1_idx, np.array(g["l1_idx"], dtype=bool))
    Y = np.random.default_rng(1).standard_normal((2, 5))
    assert np.isclose(reg.value(Y), np.sum(np.abs(reg.l1_idx * Y)))
    assert np.allclose(reg.grad(np.ones((2, 5))), np.array(g["grad_at_ones"], dtype=float))


def test_group_ard_batcharray_composite_regs():
    rng = np.random.default_rng(2)
    Y = rng.standard_normal((3, 6))
    reg = O.GroupRegularizer(G["group_regularizer"]["data_groups"], K=3)
    assert np.isclose(reg.value(Y), 0.5 * np.sum(Y ** 2)) and np.allclose(reg.grad(Y), Y)
    a = G["ard_regularizer"]
    Y = rng.standard_normal((3, 5))
    ard = O.ARDRegularizer(a["groups"])
    b = 1 + (0.5 / ard.beta[0]) * Y * Y
    assert np.isclose(ard.value(Y), (0.5 + ard.alpha[0]) * np.sum(np.log(b)))
    assert np.allclose(ard.grad(Y), ((0.5 + ard.alpha[0]) / ard.beta[0]) * Y / b)
    assert np.isclose(ard.alpha[0], a["alpha"]) and np.isclose(ard.beta[0], a["beta"])
    ba = _ba(G["batch_array_reg"])
    br = O.BatchArrayReg(ba, weight=1.0)
    assert np.isclose(br.value(ba), 0.5 * sum(np.sum(w[:, None] * v * v) for w, v in zip(br.weights, ba.values)))
    for gr, w, v in zip(br.grad(ba), br.weights, ba.values):
        assert np.allclose(gr, w[:, None] * v)
    s = G["selective_l1"]
    Y = rng.standard_normal((2, 5))
    l1 = O.SelectiveL1Reg(s["data_features"], s["edgelists"])
    net = O.NetworkRegularizer(s["data_features"], s["edgelists"])
    comp = O.CompositeRegularizer([l1, net], [0.5, 0.5])
    assert np.isclose(comp.value(Y), 0.5 * (l1.value(Y) + net.value(Y)))
    v, gsum = O._value_grad(comp, Y)
    assert np.isclose(v, comp.value(Y)) and np.allclose(gsum, 0.5 * (l1.grad(Y) + net.grad(Y)))


def test_construct_regs_mixture_weights():
    """src/regularizers.jl:655-739: 3 slots, mixture_p normalised over the enabled ones."""
    fv = [1] * 3 + [2] * 2
    el = G["selective_l1"]["edgelists"]
    yr = O.construct_Y_reg(2, 5, [1, 2, 3, 4, 5], fv, None, el, 1.0, 1.0, 1.0, False, False, None, 1.001, 0.8)
    assert len(yr.regularizers) == 3 and np.allclose(yr.mixture_p, [1 / 3] * 3)
    yr = O.construct_Y_reg(2, 5, [1, 2, 3, 4, 5], fv, None, None, 1.0, None, None, False, False, None, 1.001, 0.8)
    assert np.allclose(yr.mixture_p, [1, 0, 0])
    xr = O.construct_X_reg(2, 4, [1, 2, 3, 4], [1, 1, 2, 2], None, None, 1.0, 1.0, False, False)
    assert np.allclose(xr.mixture_p, [0, 1, 0]) and isinstance(xr.regularizers[1], O.GroupRegularizer)
    assert isinstance(O.construct_X_reg(2, 4, None, [1, 1, 2, 2], None, None, 1.0, 1.0, True, False), O.GroupRegularizer)
    assert isinstance(O.construct_X_reg(2, 4, None, None, None, None, 1.0, 1.0, False, True), O.L2Regularizer)
    assert isinstance(O.construct_Y_reg(2, 5, None, fv, None, None, 1.0, None, None, True, False, None, 1.001, 0.8),
                      O.ARDRegularizer)


def test_featureset_ard():
    g = G["featureset_ard"]
    N, K = g["N"], g["K"]
    fids = list(range(1, N + 1))
    views = [1] * 20 + [2] * 20
    reg = O.construct_featureset_ard(K, fids, views, g["feature_sets"], alpha0=g["alpha0"], lr=0.1, v0=g["v0"],
                                     dtype=np.float64)
    assert len(reg.col_ranges) == 2 and reg.featureset_ids == [[1, 2, 3, 4], [1, 2, 3, 4]]
    assert reg.alpha0 == np.float32(g["alpha0"]) and reg.v0 == np.float32(g["v0"])
    assert np.allclose(reg.beta, np.float32(g["alpha0"]) - np.float32(1))
    for i, fs in enumerate(g["feature_sets"]):
        S = np.zeros((len(fs), 20))
        for l, s in enumerate(fs):
            S[l, np.array(s) - 1 - i * 20] = 1 / np.sqrt(len(s))
        assert np.allclose(reg.S[i].toarray(), S)
        assert reg.A[i].shape == (len(fs), K)
    rng = np.random.default_rng(5)
    Y = rng.standard_normal((K, N)) * 0.3

    def gnl(beta, Y_):
        return -np.sum(reg.alpha[None, :] * np.log(beta)) + np.sum(
            (reg.alpha + 0.5)[None, :] * np.log(beta + 0.5 * Y_ * Y_))
    assert np.isclose(reg.value(Y), gnl(reg.beta, Y) - gnl(reg.beta, np.zeros_like(Y)))
    # grad == d/dY of the un-calibrated formula
    eps = 1e-6
    gr = reg.grad(Y)
    for idx in [(0, 0), (3, 17), (9, 39)]:
        Yp, Ym = Y.copy(), Y.copy()
        Yp[idx] += eps
        Ym[idx] -= eps
        assert np.isclose(gr[idx], (gnl(reg.beta, Yp) - gnl(reg.beta, Ym)) / (2 * eps), rtol=1e-5)
    # gamma_normal_loss gradient wrt A matches finite differences; update_A! runs and improves
    A = np.abs(rng.standard_normal(reg.A[0].shape)) * 0.1
    Yv = Y[:, :20]
    gA = O.gamma_normal_grad_A(A, reg.S[0], reg.alpha[:20], reg.alpha0, reg.v0, Yv)
    for idx in [(0, 0), (2, 5)]:
        Ap, Am = A.copy(), A.copy()
        Ap[idx] += eps
        Am[idx] -= eps
        fdv = (O.gamma_normal_loss(Ap, reg.S[0], reg.alpha[:20], reg.alpha0, reg.v0, Yv)
               - O.gamma_normal_loss(Am, reg.S[0], reg.alpha[:20], reg.alpha0, reg.v0, Yv)) / (2 * eps)
        assert np.isclose(gA[idx], fdv, rtol=1e-4)
    O.update_lambda(reg, Y)
    res = O.update_A(reg, Y, max_epochs=200, term_iter=20)
    assert len(res) == 2 and all(np.isfinite(r[0]) for r in res)
    assert np.all(reg.A[0] >= 0) and np.all(reg.beta > 0)


def test_ista_and_adagrad_rules():
    """src/optimizers.jl:6-13, 46-62."""
    p = np.array([[0.5, -0.2], [0.05, 1.0]], dtype=np.float32)
    g = np.array([[0.1, 0.1], [1.0, -0.5]], dtype=np.float32)
    opt = O.ISTAOptimiser(p, 0.1, np.array([2.0, 0.5], dtype=np.float32))
    ssq = np.float32(1e-8) + g * g
    eta = np.float32(0.1) / np.sqrt(ssq)
    e = np.maximum(p - eta * g, 0)
    e = np.maximum(np.abs(e) - np.array([2.0, 0.5], dtype=np.float32)[None, :] * eta, 0)
    opt.update(p, g)
    assert np.allclose(p, e)
    ad = O.AdaGrad(1.0)
    q = np.ones(3)
    gq = np.array([1.0, -2.0, 0.0])
    ad.apply("q", q, gq)
    acc = 1e-8 + gq * gq
    assert np.allclose(q, 1 - gq / (np.sqrt(acc) + 1e-8))


def test_fit_loop_contract():
    m, D, _ = O.simulate_model(30, {"methylation": ("normal", 20)}, K=3, seed=7)
    m.X_reg = O.L2Regularizer(3, 1.0)
    m.Y_reg = O.GroupRegularizer(["methylation"] * 20, K=3)
    hist = O.mf_fit_adapt_lr(m, D, lr=1.0, min_lr=0.01, max_epochs=60, update_X=True, update_Y=True)
    assert hist[-1]["term_code"] in ("max_epochs", "abs_tol", "rel_tol", "loss_increase")
    first, last = hist[0]["loss"][0], hist[-1]["loss"][-1]
    assert last < first
    for a, b in zip(hist[:-1], hist[1:]):
        assert a["term_code"] == "loss_increase" and np.isclose(b["lr"], a["lr"] * 0.5)


def test_m_estimates_closed_forms():
    """compute_M_estimates (src/fit.jl:88-92, restated): the shift that minimises a column's noise-model loss has a
    closed form for the three exponential-family models -- mean, logit(mean), log(mean) of the observed entries --
    and the restated AdaGrad loop converges to it."""
    om, D, meta = O.simulate_model(200, {"mutation": ("bernoulli", 20), "methylation": ("normal", 30), "counts": ("poisson", 15)},
                                   3, 41, batch_views=["methylation"], n_batches=3, missing=0.2)
    mu_before = om.mu.copy()
    est, h = O.compute_M_estimates(om, D, lr=0.5, max_epochs=3000, rel_tol=0.0, abs_tol=0.0)
    assert np.array_equal(om.mu, mu_before)                     # the model itself is untouched
    m = np.nanmean(D, axis=0)
    assert np.allclose(est[20:50], m[20:50], atol=1e-6)
    ok = (m[:20] > 0.02) & (m[:20] < 0.98)
    assert np.allclose(est[:20][ok], np.log(m[:20][ok] / (1 - m[:20][ok])), atol=1e-4)
    assert np.allclose(est[50:], np.log(m[50:]), atol=1e-6)
    assert h["loss"][-1] <= h["loss"][0]
    O.init_mu(om, D, lr_mu=0.5, max_epochs=50)
    assert not np.array_equal(om.mu, mu_before)


def test_ordinal_threshold_gradients_finite_differences():
    """d loss / d(t1, t2) of both ordinal noise models (update_noise_models, src/fit.jl:14) against
    central differences of the oracle's own loss."""
    m, D, meta = O.simulate_model(60, {"cna": ("ordinal3", 30), "methylation": ("ordinal_sq_hinge3", 25)}, K=3, seed=4, missing=0.1)
    g = O.data_loss_grads(m, D)
    for r in range(2):
        for c in range(2):
            eps = 1e-6
            th = m.noise.thresholds[r]
            th[1 + c] += eps
            lp = O.data_loss_grads(m, D, want_grads=False)["loss"]
            th[1 + c] -= 2 * eps
            lm = O.data_loss_grads(m, D, want_grads=False)["loss"]
            th[1 + c] += eps
            assert abs(g["dthresholds"][r, c] - (lp - lm) / (2 * eps)) <= 1e-6 * abs(g["dthresholds"][r, c]) + 1e-6


def test_alternating_and_threshold_training_are_options_of_the_oracle():
    m, D, meta = O.simulate_model(50, {"methylation": ("normal", 25), "cna": ("ordinal3", 20)}, K=3, seed=5)
    import copy
    kw = dict(max_epochs=5, update_X=True, update_Y=True, update_col_layers=True, rel_tol=0, abs_tol=0)
    runs = {}
    for name, opts in (("plain", {}), ("thr", dict(update_noise_models=True)), ("alt", dict(alternating=True))):
        mm = copy.deepcopy(m)
        h = O.mf_fit(mm, D, O.AdaGrad(0.2), **kw, **opts)
        runs[name] = (h["loss"], mm)
    assert runs["plain"][0][0] == runs["thr"][0][0] == runs["alt"][0][0]          # same first loss
    assert runs["plain"][0][-1] != runs["thr"][0][-1] and runs["plain"][0][-1] != runs["alt"][0][-1]
    assert np.array_equal(runs["plain"][1].noise.thresholds[1], m.noise.thresholds[1])
    assert not np.array_equal(runs["thr"][1].noise.thresholds[1][1:3], m.noise.thresholds[1][1:3])
    assert np.isinf(runs["thr"][1].noise.thresholds[1][[0, 3]]).all()


def test_pinning_tool_identifies_the_epoch_order(tmp_path):
    """oracle/pin_against_julia.py end to end with a stand-in for the Julia run: the "reference outputs" are written by
    the oracle with `alternating=True`; the tool must rebuild the same model from the exported files, single out that
    option combination and reject the other one.  (What it will do with julia/run_reference_fit.jl's real outputs.)"""
    import pathmatfac_b200 as P
    from pathmatfac_b200.simulate import export_problem
    from oracle import pin_against_julia as pin
    from tests.helpers import make_pair
    views = {"mutation": ("bernoulli", 12), "methylation": ("normal", 20), "mrnaseq": ("normal", 18)}
    model, om, D = make_pair(40, views, K=4, seed=11, batch_views=["methylation"], n_batches=3, n_conditions=2, missing=0.2)
    d = str(tmp_path / "exp")
    export_problem(model, d)
    m0, D0 = pin.oracle_model_from_export(d, 4)
    assert np.array_equal(np.isnan(D0), np.isnan(D)) and np.allclose(np.nan_to_num(D0), np.nan_to_num(D))
    ref = O.total_loss_grads(om, D)                     # the rebuilt model is the model the pair was made from
    got = O.total_loss_grads(m0, D0)
    assert abs(got["loss"] - ref["loss"]) <= 1e-9 * abs(ref["loss"]) and np.allclose(got["dY"], ref["dY"], rtol=1e-9, atol=1e-12)
    # stand-in for step 2: "the reference" alternates
    O.mf_fit(m0, D0, O.AdaGrad(0.05), max_epochs=4, rel_tol=0.0, abs_tol=0.0, update_X=True, update_Y=True,
             update_col_layers=True, alternating=True)
    for k, v in pin.oracle_outputs(m0).items():
        name = "ref_" + k.replace("/", "__") + ".bin"
        np.asfortranarray(v.astype(np.float32)).tofile(os.path.join(d, name)) if v.ndim < 2 else \
            open(os.path.join(d, name), "wb").write(np.asfortranarray(v.astype(np.float32)).tobytes(order="F"))
    rows, match = pin.pin(d, 4, 0.05, 4)
    assert match == {"alternating": True, "update_noise_models": False}
    assert rows[0][2] < 1e-6 and rows[1][0]["alternating"] is False and rows[1][2] > 1e-4
    assert pin.main([d, "--K", "4", "--lr", "0.05", "--epochs", "4"]) == 0
    assert pin.main([d, "--K", "4", "--lr", "0.05", "--epochs", "3"]) == 1      # a different run does not match


def _fd_grad(f, X, idxs, eps=1e-6):
    out = []
    for idx in idxs:
        old = X[idx]
        X[idx] = old + eps
        lp = f(X)
        X[idx] = old - eps
        lm = f(X)
        X[idx] = old
        out.append((lp - lm) / (2 * eps))
    return np.array(out)


def test_every_rrule_of_the_oracle_is_the_derivative_of_its_value():
    """The reference writes its pullbacks by hand (src/regularizers.jl, src/featureset_ard.jl:141-150); apart from the
    documented ColScale quirk each one is the true derivative of the value it comes with.  Central differences in
    float64 on every regulariser the constructor can produce -- including the NetworkRegularizer, whose gradient
    AA y + AB u is the derivative of the Schur-complement loss only because u minimises it (src/regularizers.jl:279-299)
    -- and on all six noise models (MatFac.jl, restated: Appendix B)."""
    rng = np.random.default_rng(17)
    K, N = 3, 11
    Y = rng.standard_normal((K, N))
    idxs = [(0, 0), (1, 4), (2, 10), (0, 7)]
    views = ["a"] * 4 + ["b"] * 7
    fids = [f"f{j}" for j in range(N)]
    edgelists = [[[fids[0], fids[1], 1.0], [fids[1], fids[2], -1.0], [fids[2], "virt", 1.0], ["virt", fids[5], 1.0],
                  [fids[6], fids[7], 1.0]] for _ in range(K)]
    fs = O.FeatureSetARDReg(K, views, [np.abs(rng.standard_normal((2, 4))), np.abs(rng.standard_normal((3, 7)))],
                            [["s1", "s2"], ["t1", "t2", "t3"]], dtype=np.float64)
    fs.beta = 0.5 + rng.random((K, N))
    fs.alpha = 1.0 + rng.random(N)
    regs = {
        "l2": O.L2Regularizer(K, 0.8),
        "group": O.GroupRegularizer(views, weight=0.7, K=K),
        "selective_l1": O.SelectiveL1Reg(fids, edgelists, weight=0.3),
        "network": O.NetworkRegularizer(fids, edgelists, epsilon=0.1, weight=1.3),
        "ard": O.ARDRegularizer(views, alpha=1.001, beta=0.4),
        "fsard": fs,
    }
    regs["composite"] = O.CompositeRegularizer([regs["l2"], regs["network"], regs["selective_l1"]], [0.5, 0.25, 0.25])
    for name, r in regs.items():
        g = r.grad(Y.copy())
        fd = _fd_grad(lambda Z: r.value(Z), Y.copy(), idxs)
        assert np.allclose([g[i] for i in idxs], fd, rtol=2e-5, atol=1e-7), (name, [g[i] for i in idxs], fd)
    # layer penalties: per-view quadratics around a centre
    v = rng.standard_normal(N)
    cp = O.ColParamReg(views, weight=0.8, center=0.3)
    assert np.allclose(cp.grad(v.copy())[[0, 5, 10]], _fd_grad(lambda z: cp.value(z), v.copy(), [0, 5, 10]), rtol=1e-5, atol=1e-8)
    # noise models: d loss / dz per entry, missing entries contribute nothing
    z = rng.standard_normal((6, 5)) * 1.5
    thr = np.array([-np.inf, -0.7, 0.9, np.inf])
    data = {"normal": rng.standard_normal((6, 5)), "bernoulli": (rng.random((6, 5)) < 0.5).astype(float),
            "poisson": rng.poisson(2.0, (6, 5)).astype(float), "bernoulli_sq_hinge": (rng.random((6, 5)) < 0.5).astype(float),
            "ordinal3": rng.integers(1, 4, (6, 5)).astype(float), "ordinal_sq_hinge3": rng.integers(1, 4, (6, 5)).astype(float)}
    for dist, a in data.items():
        a[1, 2] = np.nan
        l, g = O.noise_loss_grad(dist, z, a, thr)
        assert l[1, 2] == 0.0 and g[1, 2] == 0.0
        eps = 1e-6
        lp, _ = O.noise_loss_grad(dist, z + eps, a, thr)
        lm, _ = O.noise_loss_grad(dist, z - eps, a, thr)
        assert np.allclose(g, (lp - lm) / (2 * eps), rtol=1e-5, atol=1e-6), dist
        if dist.startswith("ordinal"):                       # and the interior thresholds (D7)
            g1, g2 = O.noise_threshold_grads(dist, z, a, thr)
            for j, gj in ((1, g1), (2, g2)):
                tp, tm = thr.copy(), thr.copy()
                tp[j] += eps
                tm[j] -= eps
                fdj = (O.noise_loss_grad(dist, z, a, tp)[0] - O.noise_loss_grad(dist, z, a, tm)[0]) / (2 * eps)
                assert np.allclose(gj, fdj, rtol=1e-5, atol=1e-6), (dist, j)
