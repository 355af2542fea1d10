"""CPU oracle for the BAMP / SCAMP / VAMP iteration loops -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's detector hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; the shipped detectors never
do (they call the CUDA library and fail loudly when it is missing).

What it follows (reference file:line, all under /root/reference):
  * ``sm_denoiser``   -> bamp.py:66-77 (mean + variance, tau halved), scamp.py:61-68 (mean only, tau halved),
                         vamp.py:96-119 (scalar tau, NOT halved)
  * ``iid_denoiser``  -> bamp.py:79-101 (``random`` mode: i.i.d. prior P0 delta_0 + Ps sum_k delta_{s_k}, no shift)
  * ``bamp_detect``   -> bamp.py:12-25 (state), 59-64 (one iteration), 116-143 (loop + allclose exit)
  * ``scamp_detect``  -> scamp.py:8-25, 43-59, 77-108
  * ``vamp_detect``   -> vamp.py:12-28, 66-94, 159-191
  * ``vamp2_detect``  -> vamp2.py:12-26, 52-87, 117-127 (the damped direct form no driver of the reference imports)

Frame semantics: the reference couples the frames of a batch (batch-global soft-max shift bamp.py:70, batch-global
exit bamp.py:140, batch-pooled variance vamp.py:85) and is numerically unusable for B>1, so a
"frame" here is ONE reference call with ``batch=1``: per-frame shift, per-frame exit, per-frame pooled variance.
The functions are vectorised over frames but every frame evolves exactly as its own batch=1 call would.

Arithmetic follows the reference's dtypes: mat-vecs and state in complex64/float32, the denoiser internals in
float64 on the complex64-rounded ``s/tau`` (symbols are complex128 in the reference, bamp.py:41), results rounded
back to complex64/float32 every iteration.  Summation order inside BLAS differs between libraries, so agreement
with the reference is to rounding (1e-6 relative per iteration), not bit-exact.

Parity pinning: the reference ships no tests or golden vectors for this path.  This oracle is
pinned against outputs of the reference itself, run in the build container on fixed-seed inputs and committed
under ``tests/golden`` by ``tests/golden/make_golden.py`` (see tests/test_oracle_golden.py).
"""
import numpy as np

F32 = np.float32
C64 = np.complex64

# torch.allclose defaults (bamp.py:140)
RTOL, ATOL = F32(1e-5), F32(1e-8)
# vamp.py:51-54
VAR_RATIO_MIN = F32(1.0e-5)
VAR_RATIO_MAX = F32(1.0) - F32(1.0e-5)
VAR_MIN, VAR_MAX = F32(1.0e-9), F32(1.0e5)


def _allclose_rows(new, old):
    """Per-frame torch.allclose(new, old): all |new-old| <= atol + rtol*|old|; NaN never closes."""
    with np.errstate(invalid='ignore'):
        ok = np.abs(new - old) <= ATOL + RTOL * np.abs(old)
    return ok.reshape(ok.shape[0], -1).all(axis=1)


def sm_denoiser(s, tau, symbols, L, M, halve_tau, shift='reference', want_var=True):
    """Section-wise spatial-modulation posterior mean (and variance).

    s: (F, L*M) complex64; tau: broadcastable to (F, L*M) float32; symbols: (K,) complex128.
    shift='reference' subtracts the frame-global max|x| (bamp.py:70, can yield NaN for L>1 at high SNR,
    SURVEY.md App. B.2); shift='section' subtracts each section's own maximum (finite everywhere, identical
    wherever the reference is finite).
    """
    F = s.shape[0]
    K = symbols.shape[0]
    s4 = np.ascontiguousarray(s, dtype=C64).reshape(F, L, M, 1)
    tau4 = np.broadcast_to(np.asarray(tau, dtype=F32).reshape(F, -1), (F, L * M)).reshape(F, L, M, 1)
    if halve_tau:
        tau4 = tau4 / F32(2)
    with np.errstate(all='ignore'):
        q = (s4 / tau4).astype(C64)                                   # complex64 division, as the reference
        sym = symbols.astype(np.complex128).reshape(1, 1, 1, K)
        x = (q.astype(np.complex128) * sym.conj()).real               # float64 from here on
        if shift == 'reference':
            ref = np.abs(x).reshape(F, -1).max(axis=1).reshape(F, 1, 1, 1)
        else:
            ref = x.reshape(F, L, M * K).max(axis=2).reshape(F, L, 1, 1)
        eta = np.exp(x - ref)
        per_antenna = eta.sum(axis=-1)                                 # (F, L, M)
        norm = per_antenna.sum(axis=2, keepdims=True)                  # (F, L, 1)
        xmmse = (sym * eta).sum(axis=-1) / norm
        out_mean = xmmse.reshape(F, L * M).astype(C64)
        if not want_var:
            return out_mean
        var0 = np.abs(xmmse) ** 2 * (1 - per_antenna / norm)
        spread = (np.abs(xmmse[..., None] - sym) ** 2 * eta).sum(axis=-1) / norm
        var = var0 + spread
    return out_mean, var.reshape(F, L * M).astype(F32)


def iid_denoiser(r, cov, symbols, sparsity):
    """``BAMPLayer.random_denoiser`` (bamp.py:79-101): posterior mean / variance under the i.i.d. prior
    ``P0 delta_0 + Ps sum_k delta_{s_k}`` with ``G(s) = exp(-|r - s|^2 / cov)`` evaluated in float64 WITHOUT any shift
    (an exact-zero normaliser is replaced by 1e-9, ``regularize_zero``).  ``Ps`` / ``P0`` are float32 tensors in the
    reference (bamp.py:40 <- config.py:86-114: Ps = sparsity / K for modulated alphabets, K = 1 for OOK)."""
    sym = np.asarray(symbols, dtype=np.complex128)
    Ps = np.float64(F32(sparsity / len(sym)))
    P0 = np.float64(F32(1.0 - sparsity))
    r128 = np.ascontiguousarray(r, dtype=C64).astype(np.complex128)[..., None]
    cov64 = np.asarray(cov, dtype=F32).astype(np.float64)[..., None]
    with np.errstate(all='ignore'):
        G0 = np.exp(-np.abs(r128) ** 2 / cov64)                              # bamp.py:91-92
        Gs = np.exp(-np.abs(r128 - sym) ** 2 / cov64)
        norm = P0 * G0 + Ps * Gs.sum(axis=-1, keepdims=True)                 # bamp.py:93
        norm = np.where(norm == 0.0, 1e-9, norm)                             # regularize_zero (bamp.py:99-101)
        ex = Ps * (sym * Gs).sum(axis=-1, keepdims=True) / norm              # bamp.py:94
        var = Ps * (np.abs(sym) ** 2 * Gs).sum(axis=-1, keepdims=True) / norm - np.abs(ex) ** 2   # bamp.py:95
    return ex[..., 0].astype(C64), var[..., 0].astype(F32)


def _mse(xmmse, x_true):
    d = xmmse - x_true
    return (d.real.astype(np.float64) ** 2 + d.imag.astype(np.float64) ** 2).mean(axis=1)


def bamp_detect(H, y, sigma2, symbols, L, M, max_iters, early_exit=True, shift='reference', x_true=None, iid_sparsity=None):
    """BAMP over F frames.  H: (n,N) shared or (F,n,N) per frame, complex64; y: (F,n) complex64.

    Returns dict(xmap, xmmse, var, iters, traj) -- ``traj`` holds per-iteration per-frame
    ``tau`` (mean effective noise variance), ``var`` (mean posterior variance) and, when ``x_true`` is given,
    ``mse``; entries of frames that already exited repeat their last value.  ``iid_sparsity`` (= Na/Nt) selects the
    ``random``-mode denoiser (bamp.py:46,79-97) instead of the sectioned one.
    """
    y = np.ascontiguousarray(y, dtype=C64)
    F, n = y.shape
    H = np.ascontiguousarray(H, dtype=C64)
    shared = H.ndim == 2
    N = H.shape[-1]
    P = (np.abs(H) ** 2).astype(F32)                                   # bamp.py:18
    Hh = np.conj(np.swapaxes(H, -1, -2))
    Pt = np.swapaxes(P, -1, -2)
    sigma2 = np.broadcast_to(np.asarray(sigma2, dtype=F32).reshape(-1, 1), (F, 1))

    xmmse = np.zeros((F, N), C64)
    var = np.ones((F, N), F32)
    z = y.copy()
    u = np.zeros((F, n), F32) + sigma2                                 # v=0 at start (bamp.py:22,25)
    xmap = np.zeros((F, N), C64)
    cov = np.zeros((F, N), F32)
    iters = np.zeros(F, np.int32)
    active = np.arange(F)
    traj = {k: np.full((max_iters, F), np.nan) for k in ('tau', 'var', 'mse')}

    def mv(Mat, vec, idx):
        if shared:
            return vec @ Mat.T
        return np.matmul(Mat[idx], vec[..., None])[..., 0]

    for t in range(max_iters):
        a = active
        if a.size == 0:
            break
        with np.errstate(all='ignore'):
            v = mv(P, var[a], a).astype(F32)                           # bamp.py:59
            resid = y[a] - z[a]
            z_new = (mv(H, xmmse[a], a) - v * resid / u[a]).astype(C64)  # bamp.py:60 (old u, new v)
            u_new = (v + sigma2[a]).astype(F32)                        # bamp.py:61
            cov_a = (F32(1) / mv(Pt, (F32(1) / u_new).astype(F32), a)).astype(F32)   # bamp.py:62
            g = ((y[a] - z_new) / u_new).astype(C64)
            xmap_a = (xmmse[a] + cov_a * mv(Hh, g, a)).astype(C64)     # bamp.py:63
            if iid_sparsity is None:
                xm, vr = sm_denoiser(xmap_a, cov_a, symbols, L, M, True, shift)   # bamp.py:64
            else:
                xm, vr = iid_denoiser(xmap_a, cov_a, symbols, iid_sparsity)
        done = _allclose_rows(vr, var[a])
        z[a], u[a], xmap[a], cov[a], xmmse[a], var[a] = z_new, u_new, xmap_a, cov_a, xm, vr
        iters[a] = t + 1
        traj['tau'][t:, a] = cov_a.mean(axis=1, dtype=np.float64)
        traj['var'][t:, a] = vr.mean(axis=1, dtype=np.float64)
        if x_true is not None:
            traj['mse'][t:, a] = _mse(xm, x_true[a])
        if early_exit:
            active = a[~done]
    return dict(xmap=xmap, xmmse=xmmse, var=var, cov=cov, iters=iters, traj=traj)


def scamp_detect(W, A, y, sigma2, symbols, cfg, max_iters, early_exit=True, shift='reference', x_true=None, psi_order='pairwise'):
    """SCAMP over F frames with a shared design matrix.  W: (Lr,Lc) float32; A: (n,N) complex64; y: (F,n).

    ``cfg`` supplies Na, Nt (=Mc), Nr (=Mr), Lin (=Lc), Lout (=Lr).  ``psi_order='reversed'`` sums ``|x|^2`` of a block
    in the opposite order (float32): psi = 1 - sum/Na is a difference of numbers near 1 and the exit test (scamp.py:105)
    compares it at 1e-8 + 1e-5 psi, so the exit iteration depends on the summation order -- the tests use the two orders to
    measure how far the reference is from itself there.
    """
    Na, Mc, Mr, Lc, Lr = cfg['Na'], cfg['Nt'], cfg['Nr'], cfg['Lin'], cfg['Lout']
    M, L = Mc // Na, Na * Lc
    W = np.ascontiguousarray(W, dtype=F32)
    A = np.ascontiguousarray(A, dtype=C64)
    Ah = np.conj(A.T)
    y = np.ascontiguousarray(y, dtype=C64)
    F, n = y.shape
    N = A.shape[1]
    sigma2 = np.broadcast_to(np.asarray(sigma2, dtype=F32).reshape(-1, 1), (F, 1))

    z = y.copy()
    psi = np.ones((F, Lc), F32)
    phi = np.full((F, Lr), np.inf, F32)
    xmmse = np.zeros((F, N), C64)
    xmap = np.zeros((F, N), C64)
    iters = np.zeros(F, np.int32)
    active = np.arange(F)
    traj = {k: np.full((max_iters, F), np.nan) for k in ('tau', 'psi', 'mse')}
    for t in range(max_iters):
        a = active
        if a.size == 0:
            break
        with np.errstate(all='ignore'):
            gma = ((psi[a] @ W.T) / F32(Lc)).astype(F32)               # scamp.py:45
            b = (gma / phi[a]).astype(F32)                             # scamp.py:47
            z_new = (y[a] - xmmse[a] @ A.T + np.repeat(b, Mr, axis=1) * z[a]).astype(C64)   # scamp.py:48
            phi_new = (sigma2[a] + gma).astype(F32)                    # scamp.py:50
            tau = (F32(L) / ((F32(1) / phi_new) @ W) / F32(Mr)).astype(F32)                 # scamp.py:52
            tau_use = np.repeat(tau, Mc, axis=1)
            phi_use = np.repeat(phi_new, Mr, axis=1)
            xmap_a = (xmmse[a] + tau_use * ((z_new / phi_use).astype(C64) @ Ah.T)).astype(C64)   # scamp.py:56
            xm = sm_denoiser(xmap_a, tau_use, symbols, L, M, True, shift, want_var=False)   # scamp.py:57
            e2 = (np.abs(xm) ** 2).astype(F32).reshape(-1, Lc, Mc)
            if psi_order == 'reversed':
                e2 = np.ascontiguousarray(e2[..., ::-1])
            p = e2.sum(axis=-1, dtype=F32)
            psi_new = (F32(1) - p / F32(Na)).astype(F32)               # scamp.py:59
        done = _allclose_rows(psi_new, psi[a])
        z[a], phi[a], xmap[a], xmmse[a], psi[a] = z_new, phi_new, xmap_a, xm, psi_new
        iters[a] = t + 1
        traj['tau'][t:, a] = tau.mean(axis=1, dtype=np.float64)
        traj['psi'][t:, a] = psi_new.mean(axis=1, dtype=np.float64)
        if x_true is not None:
            traj['mse'][t:, a] = _mse(xm, x_true[a])
        if early_exit:
            active = a[~done]
    return dict(xmap=xmap, xmmse=xmmse, psi=psi, iters=iters, traj=traj)


def vamp_detect(U, s, Vh, y, sigma2, sparsity, symbols, L, M, max_iters, early_exit=True,
                shift='reference', x_true=None, double=False):
    """VAMP over F frames.  U: (n,R) or (F,n,R); s: (R,) or (F,R); Vh: (R,N) or (F,R,N); y: (F,n).

    ``double=True`` follows the reference fed with complex128 factors (SURVEY.md hard part "complex128"):
    the linear stage runs in complex128 but the denoiser's outputs are still rounded to complex64/float32
    every iteration (vamp.py:119).
    """
    CT, RT = (np.complex128, np.float64) if double else (C64, F32)
    y = np.ascontiguousarray(y, dtype=CT)
    F, n = y.shape
    U = np.ascontiguousarray(U, dtype=CT)
    Vh = np.ascontiguousarray(Vh, dtype=CT)
    s = np.ascontiguousarray(s, dtype=RT)
    shared = Vh.ndim == 2
    R, N = Vh.shape[-2:]
    sF = np.broadcast_to(s.reshape(-1, R), (F, R))
    s2 = (sF ** 2).astype(RT)
    noise_var = np.broadcast_to(np.asarray(sigma2, dtype=np.float64).reshape(-1), (F,)).copy()   # python float per frame
    eta = R / N

    def mv(Mat, vec, idx, adj=False):
        if shared:
            Mm = np.conj(Mat.T) if adj else Mat
            return vec @ Mm.T
        Mm = np.conj(np.swapaxes(Mat[idx], -1, -2)) if adj else Mat[idx]
        return np.matmul(Mm, vec[..., None])[..., 0]

    all_idx = np.arange(F)
    y_tilde = (sF * mv(U, y, all_idx, adj=True)).astype(CT)            # vamp.py:22
    r = np.zeros((F, N), CT)
    var = np.ones((F, N), F32)
    r_tilde = np.full((F, N), sparsity, CT)
    s2t0 = sparsity ** 2 * (1 - sparsity) + (1 - sparsity) ** 2 * sparsity      # vamp.py:26, python float
    sigma2_tilde = np.full(F, s2t0, np.float64)                        # float64 holder; rounded to RT after it. 0
    xmmse = np.zeros((F, N), C64)
    iters = np.zeros(F, np.int32)
    active = all_idx
    traj = {k: np.full((max_iters, F), np.nan) for k in ('tau', 'var', 'mse', 'sigma2')}
    for t in range(max_iters):
        a = active
        if a.size == 0:
            break
        with np.errstate(all='ignore'):
            if t == 0:
                ratio = (noise_var[a] / sigma2_tilde[a]).astype(RT)    # float64 division, then the op's dtype
                s2t = sigma2_tilde[a].astype(RT)
            else:
                s2t = sigma2_tilde[a].astype(RT)
                ratio = (noise_var[a].astype(RT) / s2t).astype(RT)     # python float / 0-dim tensor (vamp.py:66)
            nv = noise_var[a].astype(RT)
            ratio_c = ratio[:, None]
            q = mv(Vh, r_tilde[a], a).astype(CT)                       # vamp.py:67
            scale = (RT(1) / (s2[a] + ratio_c)).astype(RT)             # vamp.py:68
            xt = (scale * (y_tilde[a] + ratio_c * q)).astype(CT)       # vamp.py:70
            var_lmmse = (scale.mean(axis=1, dtype=RT) * nv).astype(RT)  # vamp.py:71
            xt = (mv(Vh, (xt - q).astype(CT), a, adj=True) + r_tilde[a]).astype(CT)   # vamp.py:72
            xt_var = (RT(eta) * var_lmmse + RT(1 - eta) * s2t).astype(RT)             # vamp.py:73
            alpha = (xt_var / s2t).astype(RT)
            alpha = np.minimum(np.maximum(alpha, RT(VAR_RATIO_MIN)), RT(VAR_RATIO_MAX))
            al = alpha[:, None]
            r_a = ((xt - al * r_tilde[a]) / (RT(1) - al)).astype(CT)   # vamp.py:79
            sig2 = (alpha / (RT(1) - alpha) * s2t).astype(RT)
            sig2 = np.minimum(np.maximum(sig2, RT(VAR_MIN)), RT(VAR_MAX))
            # the denoiser divides in the input's complex dtype, works in float64, rounds its outputs (vamp.py:111-119)
            if double:
                xm, vr = _denoise_double(r_a, sig2, symbols, L, M, shift)
            else:
                xm, vr = sm_denoiser(r_a, sig2[:, None], symbols, L, M, False, shift)
            dxdr = (vr.mean(axis=1, dtype=F32).astype(RT) / sig2).astype(RT)          # vamp.py:85
            dxdr = np.minimum(np.maximum(dxdr, RT(VAR_RATIO_MIN)), RT(VAR_RATIO_MAX))
            norm = (RT(1) / (RT(1) - dxdr)).astype(RT)
            rt_new = ((xm.astype(CT) - dxdr[:, None] * r_a) * norm[:, None]).astype(CT)   # vamp.py:91
            s2t_new = (sig2 * dxdr * norm).astype(RT)
            s2t_new = np.minimum(np.maximum(s2t_new, RT(VAR_MIN)), RT(VAR_MAX))
        done = _allclose_rows(vr, var[a])
        r[a], xmmse[a], var[a], r_tilde[a] = r_a, xm, vr, rt_new
        sigma2_tilde[a] = s2t_new
        iters[a] = t + 1
        traj['tau'][t:, a] = sig2
        traj['sigma2'][t:, a] = s2t_new
        traj['var'][t:, a] = vr.mean(axis=1, dtype=np.float64)
        if x_true is not None:
            traj['mse'][t:, a] = _mse(xm, x_true[a])
        if early_exit:
            active = a[~done]
    return dict(xmap=r.astype(C64) if not double else r, xmmse=xmmse, var=var, iters=iters, traj=traj,
                y_tilde=y_tilde)


def sm_denoiser_v2(s, tau, symbols, L, M, shift='reference'):
    """``vamp2.VAMPLayer.segmented_denoiser`` (vamp2.py:78-87): the same soft-max as vamp.py's (scalar ``tau``, not halved,
    frame-global ``max|x|`` shift) but the variance is ``E|s|^2 - |E s|^2`` -- one difference of two nearly equal numbers
    where the other denoisers add two non-negative terms (it can round to a negative float32).  s: (F, L*M) complex64,
    tau: (F,) float32."""
    F = s.shape[0]
    K = symbols.shape[0]
    s4 = np.ascontiguousarray(s, dtype=C64).reshape(F, L, M, 1)
    tau4 = np.asarray(tau, dtype=F32).reshape(F, 1, 1, 1)
    with np.errstate(all='ignore'):
        q = (s4 / tau4).astype(C64)
        sym = symbols.astype(np.complex128).reshape(1, 1, 1, K)
        x = (q.astype(np.complex128) * sym.conj()).real
        if shift == 'reference':
            ref = np.abs(x).reshape(F, -1).max(axis=1).reshape(F, 1, 1, 1)
        else:
            ref = x.reshape(F, L, M * K).max(axis=2).reshape(F, L, 1, 1)
        eta = np.exp(x - ref)
        norm = eta.sum(axis=-1).sum(axis=2, keepdims=True)             # (F, L, 1)
        xmmse = (sym * eta).sum(axis=-1) / norm
        var = (np.abs(sym) ** 2 * eta).sum(axis=-1) / norm - np.abs(xmmse) ** 2
    return xmmse.reshape(F, L * M).astype(C64), var.reshape(F, L * M).astype(F32)


def vamp2_detect(U, s, Vh, y, sigma2, symbols, L, M, max_iters, damping=1.0, early_exit=True, shift='reference', x_true=None):
    """The reference's second VAMP (``vamp2.py``: "direct implementation of Rangan (with damping)"), which no driver of the
    reference imports; restated line by line: Tracker vamp2.py:12-26, one layer 52-76, loop and exit 117-127.  Frames are
    independent batch=1 calls as everywhere in this oracle.  U: (F,n,R), s: (F,R), Vh: (F,R,N), y: (F,n); sigma2 scalar or (F,)."""
    y = np.ascontiguousarray(y, dtype=C64)
    F, n = y.shape
    U = np.ascontiguousarray(U, dtype=C64)
    Vh = np.ascontiguousarray(Vh, dtype=C64)
    s = np.ascontiguousarray(s, dtype=F32)
    R, N = Vh.shape[-2:]
    sF = np.broadcast_to(s.reshape(-1, R), (F, R))
    s2 = (sF ** 2).astype(F32)
    noise_var = np.broadcast_to(np.asarray(sigma2, dtype=np.float64).reshape(-1), (F,)).copy()
    eta = N / R                                                         # vamp2.py:26 (python float)
    rho = float(damping)
    VMIN, VMAX = F32(1.0e-11), F32(1.0e11)                              # vamp2.py:48-49

    def mv(Mat, vec, adj=False):
        Mm = np.conj(np.swapaxes(Mat, -1, -2)) if adj else Mat
        if Mat.ndim == 2:
            return vec @ Mm.T
        return np.matmul(Mm, vec[..., None])[..., 0]

    with np.errstate(all='ignore'):
        y_tilde = (mv(U, y, adj=True) / sF.astype(C64)).astype(C64)     # vamp2.py:22: complex64 / float32 tensor
    r = np.zeros((F, N), C64)
    var = np.ones((F, N), F32)
    xmmse = np.zeros((F, N), C64)
    gamma = np.ones(F, F32)                                             # torch.tensor(1.0)
    iters = np.zeros(F, np.int32)
    active = np.arange(F)
    traj = {k: np.full((max_iters, F), np.nan) for k in ('gamma', 'var', 'mse')}
    for t in range(max_iters):
        a = active
        if a.size == 0:
            break
        with np.errstate(all='ignore'):
            xm, vr = sm_denoiser_v2(r[a], gamma[a], symbols, L, M, shift)                   # vamp2.py:61
            xd = (F32(rho) * xm + F32(1 - rho) * xmmse[a]).astype(C64)                      # vamp2.py:62 (python floats x complex64)
            alpha = (vr.mean(axis=1, dtype=F32) * gamma[a]).astype(F32)                     # vamp2.py:63
            al = alpha[:, None]
            r_tilde = ((xd - al * r[a]) / (F32(1) - al)).astype(C64)                        # vamp2.py:65
            g_tilde = (gamma[a] * (F32(1) - alpha) / alpha).astype(F32)                     # vamp2.py:66-68
            g_tilde = np.minimum(np.maximum(g_tilde, VMIN), VMAX)
            nv = noise_var[a].astype(F32)
            d = (s2[a] / (s2[a] + (nv * g_tilde)[:, None])).astype(F32)                     # vamp2.py:70
            dm = d.mean(axis=1, dtype=F32)
            g_new = (g_tilde * dm / (F32(eta) - dm)).astype(F32)                            # vamp2.py:71
            g_next = (F32(rho) * g_new + F32(1 - rho) * gamma[a]).astype(F32)               # vamp2.py:72 (73-74 are dead code)
            resid = (y_tilde[a] - mv(Vh[a] if Vh.ndim == 3 else Vh, r_tilde)).astype(C64)
            upd = mv(Vh[a] if Vh.ndim == 3 else Vh, ((d / dm[:, None]) * resid).astype(C64), adj=True)
            r_new = (r_tilde + F32(eta) * upd).astype(C64)                                  # vamp2.py:76
        done = _allclose_rows(vr, var[a])
        r[a], xmmse[a], var[a], gamma[a] = r_new, xd, vr, g_next
        iters[a] = t + 1
        traj['gamma'][t:, a] = g_next
        traj['var'][t:, a] = vr.mean(axis=1, dtype=np.float64)
        if x_true is not None:
            traj['mse'][t:, a] = _mse(xd, x_true[a])
        if early_exit:
            active = a[~done]
    return dict(xmap=r, xmmse=xmmse, var=var, iters=iters, traj=traj, y_tilde=y_tilde)


def _denoise_double(r, sig2, symbols, L, M, shift):
    """vamp.py:96-119 when ``r`` is complex128: the division happens in complex128, outputs round to c64/f32."""
    F = r.shape[0]
    K = symbols.shape[0]
    with np.errstate(all='ignore'):
        q = r.reshape(F, L, M, 1) / sig2.reshape(F, 1, 1, 1)
        sym = symbols.astype(np.complex128).reshape(1, 1, 1, K)
        x = (q * sym.conj()).real
        if shift == 'reference':
            ref = np.abs(x).reshape(F, -1).max(axis=1).reshape(F, 1, 1, 1)
        else:
            ref = x.reshape(F, L, M * K).max(axis=2).reshape(F, L, 1, 1)
        eta = np.exp(x - ref)
        per_antenna = eta.sum(axis=-1)
        norm = per_antenna.sum(axis=2, keepdims=True)
        xmmse = (sym * eta).sum(axis=-1) / norm
        var0 = np.abs(xmmse) ** 2 * (1 - per_antenna / norm)
        spread = (np.abs(xmmse[..., None] - sym) ** 2 * eta).sum(axis=-1) / norm
    return xmmse.reshape(F, L * M).astype(C64), (var0 + spread).reshape(F, L * M).astype(F32)


# ---- Shrink family (shrink.py:58-157): element-wise denoisers of the reference's ``random``-mode VAMP variant ------
_EXP_MAX = F32(np.log(np.finfo(np.float32).max))      # shrink.py:163-166 (regularize_exp): a[a >= max] = max - 1
TOL = F32(1.0e-9)                                      # shrink.py:28


def _regularize_exp(a):
    a = a.astype(F32).copy()
    a[a >= _EXP_MAX] = _EXP_MAX - F32(1)
    return a


def shrink_bayes(r, cov, symbols, P0, Ps):
    """shrink.py:77-95: posterior mean under P0 delta_0 + Ps sum_k delta_{s_k}, everything in complex64 / float32
    (the symbols are cast to complex64 here, shrink.py:26).  r: (..., ) complex64, cov broadcastable float32."""
    r = np.asarray(r, C64)[..., None]
    cov = np.asarray(cov, F32)[..., None] if np.ndim(cov) else F32(cov)
    sym = np.asarray(symbols).astype(C64)
    P0, Ps = F32(P0), F32(Ps)
    G0 = np.exp(-(np.abs(r) ** 2).astype(F32) / cov).astype(F32)
    Gs = np.exp(-(np.abs(r - sym) ** 2).astype(F32) / cov).astype(F32)
    norm = (P0 * G0 + Ps * Gs.sum(-1, keepdims=True, dtype=F32)).astype(F32)
    norm[norm == 0] = TOL                                 # regularize_zero, shrink.py:159-161
    return ((Ps * (sym * Gs).sum(-1, keepdims=True).astype(C64)) / norm).astype(C64)[..., 0]


def shrink_ook(r, cov, P0, Ps):
    """shrink.py:139-157: OOK posterior mean 1 / (1 + eta + tol) and the batch mean of its derivative."""
    re = np.asarray(r).real.astype(F32)
    cov = np.asarray(cov, F32)
    theta = np.log(F32(P0) / F32(Ps)).astype(F32)
    eta = np.exp(_regularize_exp(theta + (F32(1) - F32(2) * re) / cov)).astype(F32)
    e = (F32(1) / (F32(1) + eta + TOL)).astype(F32)
    with np.errstate(invalid='ignore', over='ignore'):
        der = np.nan_to_num((F32(2) * eta * e ** 2 / cov).astype(F32), nan=0.0)
    return e, F32(der.mean(dtype=np.float64))


def shrink_sw_ook(r, cov, L, M):
    """shrink.py:58-75: section-wise OOK denoiser -- extrinsic log-ratio of 'exactly one active entry per section'."""
    re = np.asarray(r).real.astype(F32)
    cov = np.asarray(cov, F32)
    B = re.shape[0]
    Lr = _regularize_exp(((F32(2) * re - F32(1)) / cov).reshape(B, L, M))
    eL = np.exp(Lr).astype(F32)
    with np.errstate(divide='ignore', invalid='ignore'):
        Le = -np.log(eL.sum(-1, keepdims=True, dtype=F32) - eL).astype(F32)
    eta = np.exp(_regularize_exp(Lr + Le)).astype(F32)
    E = (eta / (F32(1) + eta)).astype(F32)
    V = (E * (F32(1) - E)).astype(F32)
    return E.reshape(B, L * M).astype(C64), V.reshape(B, L * M)
