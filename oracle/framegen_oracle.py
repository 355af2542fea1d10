"""CPU oracle for the on-device frame generator (csrc/framegen.cuh) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the generator's counter layout and arithmetic:
  * ``philox4x32_10``  -> the Philox4x32-10 block function of Salmon et al. (Random123), pinned by the library's published
                          known-answer vectors in tests/test_framegen.py
  * ``box_muller``     -> two normals from two words, 24-bit uniforms in (0, 1), float32
  * ``frames``         -> channel.py:53-55 (i.i.d. CN(0, 1/Nr) entries), data.py:74-91 (one active antenna and one symbol
                          per section), channel.py:113-115 (y = H x + w), optional Kronecker roots H = Rr_root G Rt_root
The reference itself draws these with numpy's MT19937 / torch's generators in another order, so there is no reference output
to pin the DRAWS to: "parity unpinned" for the random sequence by construction; what is pinned is the block function (KAT),
the distribution (moments, tests) and that the two device paths consume the same stream.
See oracle/amp_oracle.py for who may import this package.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32-valued, key: (2,) -> (..., 4) uint32 (uint64 arithmetic inside)."""
    c = np.asarray(ctr, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = np.uint64(int(key[0]) & MASK), np.uint64(int(key[1]) & MASK)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def box_muller(a, b):
    u1 = ((a >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    u2 = ((b >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    th = (np.float32(6.28318530717958647692) * u2).astype(np.float32)
    return (rad * np.cos(th)).astype(np.float32), (rad * np.sin(th)).astype(np.float32)


def frames(seed, first_frame, count, n, N, M, L, symbols, gray, h_var, sigma2, Rr_root=None, Rt_root=None, frame_base=0, rho_r=0.0,
           rho_t=0.0):
    """Frames first_frame .. first_frame + count - 1 of stream `seed`: (H (F, n, N) c64, y (F, n) c64, x (F, N) c64,
    labels (F L,) int64, flat positions (F L,) int64)."""
    key = (seed & MASK, (seed >> 32) & MASK)
    sym = np.asarray(symbols).astype(np.complex64)
    gray = np.asarray(gray, dtype=np.int64)
    K = sym.shape[0]
    h_std, noise_std = np.float32(np.sqrt(h_var / 2.0)), np.float32(np.sqrt(sigma2 / 2.0))
    Hs, ys, xs, labs, idxs = [], [], [], [], []
    for f in range(count):
        fg = first_frame + f
        flo, fhi = fg & MASK, (fg >> 32) & MASK
        nb = n * N // 2
        ctr = np.zeros((nb, 4), dtype=np.uint64)
        ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3] = flo, fhi, 0, np.arange(nb)
        w = philox4x32_10(ctr, key)
        r0, i0 = box_muller(w[:, 0], w[:, 1])
        r1, i1 = box_muller(w[:, 2], w[:, 3])
        G = np.empty(n * N, dtype=np.complex64)
        G[0::2] = (r0 * h_std) + 1j * (i0 * h_std)
        G[1::2] = (r1 * h_std) + 1j * (i1 * h_std)
        H = G.reshape(n, N)
        if rho_t:                                                   # AR(1) along the columns, then along the rows (float32 FMAs)
            H = H.copy()
            r32, c32 = np.float32(rho_t), np.float32(np.sqrt(1.0 - rho_t * rho_t))
            for c in range(1, N):
                H[:, c] = (r32 * H[:, c - 1].real + c32 * H[:, c].real) + 1j * (r32 * H[:, c - 1].imag + c32 * H[:, c].imag)
        if rho_r:
            H = H.copy()
            r32, c32 = np.float32(rho_r), np.float32(np.sqrt(1.0 - rho_r * rho_r))
            for r in range(1, n):
                H[r] = (r32 * H[r - 1].real + c32 * H[r].real) + 1j * (r32 * H[r - 1].imag + c32 * H[r].imag)
        if Rt_root is not None:
            H = (H.astype(np.complex128) @ np.asarray(Rt_root, dtype=np.complex128)).astype(np.complex64)
        if Rr_root is not None:
            H = (np.asarray(Rr_root, dtype=np.complex128) @ H.astype(np.complex128)).astype(np.complex64)
        ctr = np.zeros((L, 4), dtype=np.uint64)
        ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3] = flo, fhi, 2, np.arange(L)
        w = philox4x32_10(ctr, key).astype(np.uint64)
        pos = np.arange(L) * M + ((w[:, 0] * np.uint64(M)) >> np.uint64(32)).astype(np.int64)
        ks = ((w[:, 1] * np.uint64(K)) >> np.uint64(32)).astype(np.int64)
        x = np.zeros(N, dtype=np.complex64)
        x[pos] = sym[ks]
        nn = (n + 1) // 2
        ctr = np.zeros((nn, 4), dtype=np.uint64)
        ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3] = flo, fhi, 1, np.arange(nn)
        w = philox4x32_10(ctr, key)
        r0, i0 = box_muller(w[:, 0], w[:, 1])
        r1, i1 = box_muller(w[:, 2], w[:, 3])
        nz = np.empty(2 * nn, dtype=np.complex64)
        nz[0::2] = (r0 * noise_std) + 1j * (i0 * noise_std)
        nz[1::2] = (r1 * noise_std) + 1j * (i1 * noise_std)
        y = (H.astype(np.complex128) @ x.astype(np.complex128)).astype(np.complex64) + nz[:n]
        Hs.append(H), ys.append(y), xs.append(x), labs.append(gray[ks]), idxs.append(pos + (frame_base + f) * N)
    return (np.stack(Hs), np.stack(ys), np.stack(xs), np.concatenate(labs), np.concatenate(idxs))
