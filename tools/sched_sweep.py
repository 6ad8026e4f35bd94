"""Fused-apply time vs CTA schedule for per-rank problem sizes (one GPU emulating one slab)."""
import os, sys, subprocess
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    import perphil_b200 as pb
    from perphil_b200.solver import configure_handle

    def configured_handle(W, prm, bcs):
        h = pb.handle_for(W)
        configure_handle(h, W, prm, bcs)
        return h
    nx = int(sys.argv[1])
    mesh = pb.UnitCubeMesh(nx, 256, 256)
    _, V = pb.create_function_spaces(mesh); W = V * V
    prm = pb.DPPParameters(k1=1.0, k2=1e-2, beta=1.0, mu=1.0)
    bcs = [pb.DirichletBC(W.sub(0), pb.Constant(1.0), "on_boundary"), pb.DirichletBC(W.sub(1), pb.Constant(0.0), "on_boundary")]
    h = configured_handle(W, prm, bcs)
    a, u, m = h.time_cg_kernels(reps=20, warmup=3)
    print(f"nx={nx:4d} sched={os.environ.get('DPP_FUSED_SCHED','auto'):>4s}  fused apply {a*1e3:7.1f} us  r_update {u*1e3:6.1f} us  matvec {m*1e3:6.1f} us", flush=True)
else:
    for nx, scheds in ((32, ["auto", "p", "1", "2", "3", "4"]), (64, ["auto", "p", "2", "3", "4", "6"]), (128, ["auto", "p", "4", "6", "8"]), (256, ["auto", "7", "9", "11"])):
        for sc in scheds:
            env = dict(os.environ)
            if sc != "auto":
                env["DPP_FUSED_SCHED"] = sc
            subprocess.run([sys.executable, __file__, str(nx)], env=env)
