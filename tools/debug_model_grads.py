"""GPU: print relative gradient-norm differences of RegNet_v2 vs the reference golden, with
cuDNN on and off (the reference trains with cuDNN off, src/deterministic.py:37)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_host_logic_cpu import build_model, load_golden_model, run_model, _bias_cancelled_by_bn, _rel
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
for cudnn in (True, False):
    torch.backends.cudnn.enabled = cudnn
    g, state = load_golden_model()
    model = build_model(state, "cuda:0")
    out3, out4, loss, inter = run_model(model, g, "cuda:0")
    print("cudnn", cudnn, "out3 rel", _rel(out3.detach().cpu(), g["out3"]), "out4 rel", _rel(out4.detach().cpu(), g["out4"]),
          "loss", float(loss), float(g["loss"]))
    grads = {n: p.grad for n, p in model.named_parameters()}
    rows = []
    for n, ref in zip([str(x) for x in g["grad_names"]], g["grad_norms"]):
        if _bias_cancelled_by_bn(n): continue
        rows.append((abs(float(grads[n].norm()) - ref) / max(ref, 1e-12), n, ref))
    rows.sort(reverse=True)
    for r in rows[:12]: print("   %.3e  %-55s ref %.4g" % r)
    for k in g.files:
        if k.startswith("grad__"):
            print("   elementwise", k, _rel(grads[k[6:]].cpu(), g[k]))
