"""Monte-Carlo Eb/N0 sweep on the GPU(s) -- the driver the reference leaves empty (/root/reference/simulate.py, 0 bytes).

    python simulate.py --alg bamp --frames 1000000 --path Simulations/BAMP/run1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        simulate.py --alg vamp --channel kronecker --rho-t 0.9 --rho-r 0.9 --frames 100000000 --path Simulations/VAMP/c5

See amp-sparc-spatialmodulation_b200/simulate.py for what a sweep does (bamp_model.py:44-67 on device-generated frames).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import __graft_entry__ as ge  # noqa: E402

if __name__ == "__main__":
    ge.build()
    from amp_sparc_spatialmodulation_b200.simulate import main
    main()
