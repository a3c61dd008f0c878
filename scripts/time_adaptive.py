import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import _engine as E, solve as S
from phiml_b200 import stencils
dev = torch.device("cuda", 0)
S.RESIDENT['mode'] = 'never'
n = 256
sh = stencils.neg_laplace_pairs(3)
st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in sh], [float(c) for _, c in sh], torch.float32, dev)
y = torch.as_tensor(np.random.default_rng(0).standard_normal((1, n ** 3)).astype(np.float32), device=dev)
for method, code in (("CG (torch rules)", E.CG_TORCH), ("auto (torch rules)", E.CG_ADAPTIVE_TORCH), ("CG-adaptive (generic)", E.CG_ADAPTIVE)):
    if code == E.CG_TORCH:
        res = pb.linear_solve('CG', st, y, torch.zeros_like(y), [1e-5], [0.0], np.array([[5000]]), None, None)
    elif code == E.CG_ADAPTIVE_TORCH:
        res = pb.linear_solve('auto', st, y, torch.zeros_like(y), [1e-5], [0.0], np.array([[5000]]), None, None)
    else:
        res = pb.generic_conjugate_gradient_adaptive(st, y, torch.zeros_like(y), [1e-5], [0.0], np.array([[5000]]))
    tr = float((y - st.matvec(res.x)).norm() / y.norm())
    z = torch.zeros(1, device=dev); mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)
    solver = pb.DeviceSolver(st, None, code, 0, 1)
    solver.start(y, torch.zeros_like(y), z, z, mi); solver.advance(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); solver.advance(400); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 400
    print(json.dumps({"method": method, "fused": os.environ.get("PHB200_ADAPTIVE_FUSED", "1") + os.environ.get("PHB200_CG_TORCH_FUSED", "1"), "iterations": int(res.iterations[0]), "fev": int(res.function_evaluations[0]), "converged": bool(res.converged[0]),
                      "true_residual": tr, "us_per_iteration": ms * 1e3, "it_per_s": 1e3 / ms}))
