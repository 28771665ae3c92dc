"""Record outputs of the reference's OWN CUDA kernels on a B200 (run under gpurun).

oracle/_ref/{pointnet2_cuda,fused_conv_select_k_cuda}.so are the reference's extension sources
compiled unmodified for sm_100 (oracle/build_ref.py).  This script feeds them the seeded inputs
of tests/ref_cases.py and stores what they return in gpurun_out/ref_gpu_golden.npz; the file is
then committed as tests/golden/ref_gpu_golden.npz.  It pins the C oracle (checked against these
vectors on CPU, tests/test_oracle_golden.py) and, through it, the sm_100a kernels.

    gpurun -- python tests/golden/make_ref_gpu_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import ref_cases  # noqa: E402


def main():
    pn2, fused = ref_cases.load_reference_extensions()
    dev = torch.device("cuda:0")
    out = {}
    for name, case in ref_cases.all_cases().items():
        res = ref_cases.run_reference(case, pn2, fused, dev)
        for k, v in res.items():
            out["%s__%s" % (name, k)] = v
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ref_gpu_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.2f MB" % (os.path.getsize(path) / 1e6), len(out), "arrays")


if __name__ == "__main__":
    main()
