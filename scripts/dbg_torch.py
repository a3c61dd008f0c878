import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import phiml_b200 as pb
from phiml_b200 import _engine as E, solve as S
from oracle.assemble import laplace_shifts
dev = torch.device("cuda", 0)
S.RESIDENT['mode'] = 'never'
for n in (64, 128):
    sh = laplace_shifts(3)
    st = pb.Matrix.stencil_matrix((n, n, n), [s for s, _ in sh], [float(c) for _, c in sh], torch.float32, dev)
    y = torch.as_tensor(np.random.default_rng(0).standard_normal((1, n ** 3)).astype(np.float32), device=dev)
    res = pb.linear_solve('CG', st, y, torch.zeros_like(y), [1e-5], [0.0], np.array([[5000]]), None, None)
    tr = float((y - st.matvec(res.x)).norm() / y.norm())
    rr = float(res.residual.norm() / y.norm())
    drift = float((y - st.matvec(res.x) - res.residual).norm() / y.norm())
    # fixed number of passes: compare x after exactly 45 passes (2 refreshes) 
    z = torch.zeros(1, device=dev); mi = torch.full((1,), 2 ** 30, dtype=torch.int32, device=dev)
    solver = pb.DeviceSolver(st, None, E.CG_TORCH, 0, 1)
    solver.start(y, torch.zeros_like(y), z, z, mi); solver.advance(45); torch.cuda.synchronize()
    x45 = solver.x.clone(); r45 = solver.r.clone()
    d45 = float((y - st.matvec(x45) - r45).norm() / y.norm())
    print(json.dumps({"n": n, "fused": os.environ.get("PHB200_CG_TORCH_FUSED", "1"), "its": int(res.iterations[0]), "true_res": tr, "recurrence_res": rr, "drift |y-Ax-r|": drift,
                      "after 45 passes: |y-Ax-r|/|y|": d45, "x45_norm": float(x45.norm())}))
