"""Host logic of the fused CFG+scheduler step: the per-step linear coefficients must reproduce the
oracle's step-by-step scheduler arithmetic (oracle/restated.py, which restates the Swift twins) and
the closed-form identities available without a runnable reference (parity unpinned)."""
import numpy as np
import pytest
import torch

from b200sd import scheduler as S
from oracle import restated as R


def _run_plan(sched, eps_fn, x0, guidance=7.5, start=0):
    x = x0.copy()
    hist = [np.zeros_like(x) for _ in range(4)]
    xs, x0s = [], []
    for st in sched.plan(start=start):
        eu, ec = eps_fn(x, st.timestep)
        x, den = S.apply_plan_host(st, guidance, eu, ec, x, hist)
        xs.append(x.copy())
        x0s.append(den.copy())
    return xs, x0s


def _eps_fn(seed):
    rng = np.random.RandomState(seed)
    w = rng.randn(2, 8).astype(np.float64)

    def f(x, t):
        base = np.tanh(x * 0.7 + t / 1000.0)
        return base * w[0, 0] + 0.1, base * w[1, 1] - 0.05
    return f


def test_alphas_cumprod_matches_oracle():
    a = S.alphas_cumprod()
    b = R.alphas_cumprod().numpy()
    assert np.allclose(a, b, rtol=2e-6, atol=0)
    assert a.shape == (1000,) and 0.99 < a[0] < 1 and a[-1] < 0.01


def test_ddim_timesteps_and_identity():
    s = S.DDIMScheduler(20)
    s.abar = R.alphas_cumprod().double().numpy()  # same fp32 table on both sides
    assert s.timesteps == R.leading_timesteps(20) == list(range(951, 0, -50))
    assert S.DDIMScheduler(50).timesteps[0] == 981  # BASELINE config 1 timestep
    abar = R.alphas_cumprod().double()
    rng = np.random.RandomState(0)
    x = rng.randn(4, 8, 8)
    f = _eps_fn(1)
    xs, x0s = _run_plan(s, f, x)
    xr = torch.from_numpy(x.copy())
    for i, t in enumerate(s.timesteps):
        eu, ec = f(xr.numpy(), t)
        eps = torch.from_numpy(R.cfg_combine(eu, ec, 7.5))
        a_t = abar[t]
        x0_ref = (xr - (1 - a_t).sqrt() * eps) / a_t.sqrt()
        xr = R.ddim_step(eps, t, xr, abar, 20)
        assert np.allclose(xs[i], xr.numpy(), rtol=1e-9, atol=1e-9)
        assert np.allclose(x0s[i], x0_ref.numpy(), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("n", [20, 10, 50, 14, 15])
def test_dpm_matches_oracle(n):
    s = S.DPMSolverMultistepScheduler(n)
    s.abar = R.alphas_cumprod().double().numpy()
    ref = R.DPMSolverPP2M(n, abar=R.alphas_cumprod().double())
    assert s.timesteps == ref.timesteps
    rng = np.random.RandomState(0)
    x = rng.randn(4, 8, 8)
    f = _eps_fn(2)
    xs, x0s = _run_plan(s, f, x)
    xr = torch.from_numpy(x.copy())
    for i, t in enumerate(ref.timesteps):
        eu, ec = f(xr.numpy(), t)
        xr = ref.step(torch.from_numpy(R.cfg_combine(eu, ec, 7.5)), i, xr)
        assert np.allclose(xs[i], xr.numpy(), rtol=1e-8, atol=1e-8), (n, i)
        assert np.allclose(x0s[i], ref.x0_hist[-1].numpy(), rtol=1e-8, atol=1e-8)


def test_dpm_first_order_equals_ddim_update():
    # DPMSolverMultistepScheduler.swift:153-174 "equivalent to DDIM": with the same (t, t_prev) pair the
    # first-order DPM-Solver++ update and the DDIM eta=0 update coincide
    abar = S.alphas_cumprod().astype(np.float64)
    for t, p in [(951, 901), (501, 451), (51, 1)]:
        a, s = np.sqrt(abar), np.sqrt(1 - abar)
        lam = np.log(a) - np.log(s)
        h = lam[p] - lam[t]
        x, eps = 0.37, -1.2
        x0 = (x - s[t] * eps) / a[t]
        dpm = (s[p] / s[t]) * x - a[p] * (np.exp(-h) - 1) * x0
        ddim = a[p] * x0 + s[p] * eps
        assert abs(dpm - ddim) < 1e-12


@pytest.mark.parametrize("n", [20, 10, 50, 4])
def test_pndm_matches_oracle(n):
    s = S.PNDMScheduler(n)
    s.abar = R.alphas_cumprod().double().numpy()
    ref = R.PNDM(n, abar=R.alphas_cumprod().double())
    assert s.timesteps == ref.timesteps and len(s.timesteps) == n + 1
    if n == 20:
        assert s.timesteps[:4] == [951, 901, 901, 851]
    rng = np.random.RandomState(0)
    x = rng.randn(4, 8, 8)
    f = _eps_fn(3)
    xs, _ = _run_plan(s, f, x)
    xr = torch.from_numpy(x.copy())
    for i, t in enumerate(ref.timesteps):
        eu, ec = f(xr.numpy(), t)
        xr = ref.step(torch.from_numpy(R.cfg_combine(eu, ec, 7.5)), t, xr)
        assert np.allclose(xs[i], xr.numpy(), rtol=1e-8, atol=1e-8), (n, i)


def test_unknown_scheduler():
    with pytest.raises(ValueError):
        S.make_scheduler("Euler", 20)


def test_prepare_latents_follows_the_reference_rng_vector():
    """The reference seeds numpy globally and draws the initial latents with np.random.randn (pipeline.py:322-344,
    main :735); its Swift twin pins that stream with a known-answer test (StableDiffusionTests.swift:46-61:
    seed 12345, last five of 10 000 normals).  prepare_latents must consume the same stream in the same order."""
    import types

    from b200sd.pipeline import B200StableDiffusionPipeline

    stub = types.SimpleNamespace(vae_scale_factor=8)
    np.random.seed(12345)
    lat = B200StableDiffusionPipeline.prepare_latents(stub, 1, 4, 400, 400)  # 4 * 50 * 50 = 10 000 draws
    assert lat.shape == (1, 4, 50, 50) and lat.dtype == np.float32
    expected = np.array([-0.86285345, 2.15229409, -0.00670556, -1.21472309, 0.65498866])
    got = lat.reshape(-1)[-5:]
    assert np.allclose(got, expected.astype(np.float16).astype(np.float32), atol=0, rtol=0), got
    with pytest.raises(ValueError, match="Unexpected latents shape"):
        B200StableDiffusionPipeline.prepare_latents(stub, 1, 4, 400, 400, latents=np.zeros((1, 4, 8, 8)))


def test_image_to_image_start_step_add_noise_and_truncated_plans():
    """Scheduler.swift:83-114: startStep = max(n - Int(Float(n) * strength), 0), timesteps[startStep...],
    noisy = sqrt(abar_t) x0 + sqrt(1 - abar_t) noise at t = timesteps[startStep]."""
    from b200sd import scheduler as S

    d = S.DDIMScheduler(20)
    assert d.start_step(0.5) == 10 and d.start_step(1.0) == 0 and d.start_step(0.0) == 20 - 0
    assert d.start_step(0.75) == 5 and d.start_step(0.3) == 20 - int(np.float32(20) * np.float32(0.3))
    assert d.calculate_timesteps(0.5) == d.timesteps[10:] and d.calculate_timesteps() == d.timesteps
    x0, nz = np.full((1, 4, 2, 2), 2.0, np.float32), np.full((1, 4, 2, 2), -1.0, np.float32)
    t = d.timesteps[10]
    a = float(S.alphas_cumprod()[t])
    assert np.allclose(d.add_noise(x0, nz, 0.5), np.sqrt(a) * 2.0 - np.sqrt(1 - a), rtol=1e-6)
    assert [p.timestep for p in d.plan(start=10)] == d.timesteps[10:]
    full, cut = d.plan(), d.plan(start=10)
    assert all(abs(a_.cx - b_.cx) < 1e-15 and abs(a_.ce - b_.ce) < 1e-15 for a_, b_ in zip(full[10:], cut))
    m = S.DPMSolverMultistepScheduler(20)
    cut = m.plan(start=8)
    assert [p.timestep for p in cut] == m.timesteps[8:]
    assert cut[0].n_hist == 0 and cut[1].n_hist == 2          # the multistep state starts empty at the start step
    assert m.plan()[8].n_hist == 2                            # ... unlike step 8 of the full schedule


@pytest.mark.parametrize("n,start", [(20, 10), (20, 3), (10, 5), (14, 9)])
def test_truncated_dpm_and_pndm_plans_match_fresh_oracle_schedulers(n, start):
    """A fresh scheduler fed timeSteps[startStep...] (Scheduler.swift:109-114): the multistep state starts empty."""
    rng = np.random.RandomState(1)
    x = rng.randn(4, 8, 8)
    f = _eps_fn(2)
    # DPM-Solver++: step index into the FULL timestep list, empty history
    s = S.DPMSolverMultistepScheduler(n)
    s.abar = R.alphas_cumprod().double().numpy()
    ref = R.DPMSolverPP2M(n, abar=R.alphas_cumprod().double())
    xs, _ = _run_plan(s, f, x, start=start)
    xr = torch.from_numpy(x.copy())
    for j, i in enumerate(range(start, n)):
        eu, ec = f(xr.numpy(), ref.timesteps[i])
        xr = ref.step(torch.from_numpy(R.cfg_combine(eu, ec, 7.5)), i, xr)
        assert np.allclose(xs[j], xr.numpy(), rtol=1e-8, atol=1e-8), ("dpm", n, start, i)
    # PNDM: counter-driven branches applied to the sliced list
    p = S.PNDMScheduler(n)
    p.abar = R.alphas_cumprod().double().numpy()
    refp = R.PNDM(n, abar=R.alphas_cumprod().double())
    xs, _ = _run_plan(p, f, x, start=start)
    assert [st.timestep for st in p.plan(start=start)] == refp.timesteps[start:]
    xr = torch.from_numpy(x.copy())
    for j, t in enumerate(refp.timesteps[start:]):
        eu, ec = f(xr.numpy(), t)
        xr = refp.step(torch.from_numpy(R.cfg_combine(eu, ec, 7.5)), t, xr)
        assert np.allclose(xs[j], xr.numpy(), rtol=1e-8, atol=1e-8), ("pndm", n, start, j)


@pytest.mark.parametrize("n", [5, 14, 20])
def test_dpm_final_sigmas_zero_lands_on_the_denoised_estimate(n):
    """diffusers 0.30.2 (final_sigmas_type="zero", what the reference's Python pipeline runs, pipeline.py:565-569): the
    last step is first order with sigma_next = 0, i.e. x_prev = x0; every earlier step equals the Swift-ending plan."""
    from b200sd import scheduler as S
    zero = S.DPMSolverMultistepScheduler(n, final_sigmas_type="zero").plan()
    swift = S.DPMSolverMultistepScheduler(n).plan()
    for a, b in zip(zero[:-1], swift[:-1]):
        assert a == b
    last = zero[-1]
    assert last.cx == last.x0_cx and last.ce == last.x0_ce and last.n_hist == 0
    rng = np.random.RandomState(0)
    x, eu, ec = (rng.randn(4, 8, 8).astype(np.float32) for _ in range(3))
    hist = [np.zeros_like(x) for _ in range(4)]
    xp, x0 = S.apply_plan_host(last, 7.5, eu, ec, x, hist)
    np.testing.assert_array_equal(xp, x0)
    with pytest.raises(ValueError):
        S.DPMSolverMultistepScheduler(n, final_sigmas_type="bogus")


def test_pipeline_uses_the_diffusers_ending_for_dpm():
    from b200sd import scheduler as S
    assert S.make_scheduler("DPMSolverMultistep", 20, final_sigmas_type="zero").final_sigmas_type == "zero"


@pytest.mark.parametrize("n", [2, 3, 4, 5])
def test_pndm_first_steps_equal_the_swift_formulas_evaluated_by_hand(n):
    """Pin without the oracle: the first model evaluations of PLMS written out from Scheduler.swift:218-343.
    Step 0 is a plain formula-(9) update from t to t - d with e0.  The schedule then repeats its second timestep: step 1
    restarts from the saved first sample with (e1 + e0) / 2 over [t, t - d] and does not store e1.  Steps 2, 3, 4 use the
    2-, 3- and 4-term Adams-Bashforth weights over the stored outputs [e0, e2, e3, e4].  Guidance 1: eps = cond."""
    d = 1000 // n
    fwd = [int(round(i * float(d))) + 1 for i in range(n)]            # Scheduler.swift:188-191 (steps offset 1)
    ts = (fwd[:-1] + [fwd[-2]] + [fwd[-1]])[::-1] if n > 1 else fwd   # :197-201
    s = S.PNDMScheduler(n)
    assert s.timesteps == ts
    betas = np.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=np.float64) ** 2
    abar = np.cumprod(1.0 - betas)
    s.abar = abar

    def prev_sample(x, t, tp, e):                                      # formula (9), :305-343
        a_t, a_p = abar[t], abar[max(0, tp)]
        coeff = np.sqrt(a_p / a_t)
        denom = a_t * np.sqrt(1 - a_p) + np.sqrt(a_t * (1 - a_t) * a_p)
        return coeff * x - (a_p - a_t) / denom * e

    rng = np.random.RandomState(n)
    x0 = rng.randn(3, 5)
    es = [rng.randn(3, 5) for _ in range(n + 1)]
    k = iter(range(n + 1))
    xs, _ = _run_plan(s, lambda x, t: (np.zeros_like(x), es[next(k)]), x0, guidance=1.0)
    t0 = ts[0]
    x1 = prev_sample(x0, t0, t0 - d, es[0])                            # counter 0
    assert np.allclose(xs[0], x1, rtol=1e-12, atol=1e-12)
    # counter 1: timestep ts[1] = t0 - d arrives; prevStep = ts[1], timeStep = ts[1] + d = t0, sample = the saved x0
    x2 = prev_sample(x0, t0, ts[1], 0.5 * es[1] + 0.5 * es[0])
    assert ts[1] == t0 - d and np.allclose(xs[1], x2, rtol=1e-12, atol=1e-12)
    if n >= 2:
        # counter 2: ets = [e0, e2] (the counter-1 output is not stored): (3 e2 - e0) / 2 from ts[2] to ts[2] - d
        x3 = prev_sample(x2, ts[2], ts[2] - d, 1.5 * es[2] - 0.5 * es[0])
        assert np.allclose(xs[2], x3, rtol=1e-12, atol=1e-12)
    if n >= 3:
        x4 = prev_sample(x3, ts[3], ts[3] - d, (23 * es[3] - 16 * es[2] + 5 * es[0]) / 12.0)
        assert np.allclose(xs[3], x4, rtol=1e-12, atol=1e-12)
    if n >= 4:
        x5 = prev_sample(x4, ts[4], ts[4] - d, (55 * es[4] - 59 * es[3] + 37 * es[2] - 9 * es[0]) / 24.0)
        assert np.allclose(xs[4], x5, rtol=1e-12, atol=1e-12)
