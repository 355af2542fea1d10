"""Diagnostic: VAMP at C3 (128 x 64, Na = 4, QPSK, 2 dB) -- kernel against the oracle on the same CPU-generated frames."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from oracle import amp_oracle as ao  # noqa: E402

F = 40
c = pkg.Config(128, 4, 64, 1, 1, batch=F, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform',
               device='cpu')
np.random.seed(0)
torch.manual_seed(0)
ch, da = pkg.Channel(c), pkg.Data(c)
snr = 10 ** (2.0 / 10)
_, A = ch.generate_as_sparc()
U, s, Vh = torch.linalg.svd(A, full_matrices=False)
x, sym, i = da.generate_message()
y = A @ x + ch.awgn(snr)
r = ao.vamp_detect(np.broadcast_to(U.numpy(), (F,) + U.shape), np.broadcast_to(s.numpy(), (F,) + s.shape),
                   np.broadcast_to(Vh.numpy(), (F,) + Vh.shape), y.numpy().reshape(F, -1), (c.Na / c.Nr) / snr, c.Na / c.Nt,
                   c.symbols, c.L, c.M, 20, x_true=x.numpy().reshape(F, -1), shift='section')
print("oracle iters", r["iters"].tolist())
cg = pkg.Config(128, 4, 64, 1, 1, batch=F, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform',
                device='cuda:0')
for exp in ("f64", "f32"):
    for where in ("cpu-svd", "gpu-svd"):
        if where == "gpu-svd":
            Ug, sg, Vg = torch.linalg.svd(A.cuda(), full_matrices=False)
        else:
            Ug, sg, Vg = U.cuda(), s.cuda(), Vh.cuda()
        amp = pkg.VAMP(cg, exp=exp, trajectory=True)
        d = amp.detect(Ug, sg, Vg, y.cuda(), snr, x.cuda(), sym, i)
        cd = d.counters_dict()
        print(exp, where, "iters", d.iters.cpu().tolist(), "index_err", cd["index_err"], "nan", cd["nan_frames"])
        tr = d.traj.cpu().numpy()
        print("   frame0 sigma2_tilde", np.array2string(tr[0, :, 0], precision=4))
        print("   oracle frame0      ", np.array2string(np.asarray(r["traj"]["sigma2t"])[:, 0] if "sigma2t" in r["traj"] else np.zeros(1), precision=4))
