"""Shared comparison helpers for the parity tests (oracle vs golden, CUDA vs oracle, CUDA vs golden)."""
import numpy as np

from oracle import loss_oracle as lo

# Floors below which a trajectory value is rounding noise (SURVEY.md section 7 "Tolerance realism").
VALUE_FLOOR = 1e-9


def rel_err(a, b, floor=VALUE_FLOOR):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    mask = np.isfinite(b) & (np.abs(b) > floor)
    out = np.zeros(b.shape)
    out[mask] = np.abs(a - b)[mask] / np.abs(b)[mask]
    return out


def check_trajectory(name, got, want, tight_iters=2, tight=1e-4, median_tol=1e-4, loose=5e-2, floor=VALUE_FLOOR):
    """got/want: (frames, T).  First `tight_iters` iterations per frame within `tight` (the north-star
    tolerance for complex64); later iterations amplify float32 rounding chaotically on a few frames
    (SURVEY.md section 7), so they are held to the median and a loose per-frame cap."""
    r = rel_err(got, want, floor)
    assert r[:, :tight_iters].max() <= tight, f"{name}: early iterations off by {r[:, :tight_iters].max():.3e}"
    assert np.median(r) <= median_tol, f"{name}: median rel err {np.median(r):.3e}"
    assert r.max() <= loose, f"{name}: worst rel err {r.max():.3e}"
    return r


def counters_for(cfg, xmap, xmmse, x, sym, idx, iters=None, frames_per_call=None):
    dims = dict(Nt=cfg.Nt, Na=cfg.Na, Lin=cfg.Lin)
    return lo.error_counters(np.asarray(xmap).astype(np.complex64), xmmse, x, np.asarray(sym).ravel(), np.asarray(idx).ravel(),
                             cfg.symbols, cfg.gray, dims, iters=iters,
                             decision='sparc' if cfg.mode == 'sparc' else 'segmented')


def decision_mismatch_frames(cfg, xmap_a, xmap_b):
    """Frames whose hard decisions differ between two estimates (near-tie listing)."""
    dec = lo.map_decision if cfg.mode == 'sparc' else lo.segmented_decision
    M = cfg.Nt // cfg.Na
    F = xmap_a.shape[0]
    _, ant_a, k_a = dec(np.asarray(xmap_a).astype(np.complex64), cfg.symbols, cfg.gray, M)
    _, ant_b, k_b = dec(np.asarray(xmap_b).astype(np.complex64), cfg.symbols, cfg.gray, M)
    bad = ((ant_a != ant_b) | (k_a != k_b)).reshape(F, -1).any(axis=1)
    return np.nonzero(bad)[0]


INT_KEYS = ['frames', 'frame_err', 'slot_err', 'slot_err_first', 'slot_err_mid', 'slot_err_last',
            'index_err', 'symbol_err', 'index_bit_err', 'symbol_bit_err']


def assert_counts_equal(name, got, want, keys=INT_KEYS):
    diff = {k: (got[k], want[k]) for k in keys if int(got[k]) != int(want[k])}
    assert not diff, f"{name}: counter mismatch {diff}"
