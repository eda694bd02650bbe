import sys, torch
sys.path.insert(0, ".")
import mini_b200 as mb
for a in ("quad", "rescan"):
    ctx = mb.Context(0)
    ctx.set_advance_impl(mb.ADVANCE_QUAD if a == "quad" else mb.ADVANCE_QUAD_RESCAN)
    g = ctx.prepare_graph(ctx.rmat_graph(22, 16, 1))
    gw = ctx.rmat_graph(22, 16, 1, weighted=True)
    lab = torch.empty(g.n, dtype=torch.int32, device="cuda")
    dist = torch.empty(g.n, dtype=torch.float32, device="cuda")
    def t(fn, reps=30):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    print(a, "push %.4f" % t(lambda: ctx.bfs(g, 0, mb.BFS_PUSH, labels=lab)), "beamer %.4f" % t(lambda: ctx.bfs(g, 0, mb.BFS_BEAMER, 15.0, 18.0, labels=lab)), "sssp %.4f" % t(lambda: ctx.sssp(gw, 0, dist=dist), 10))
    acc = None
    for _ in range(5):
        _, st = ctx.bfs(g, 0, mb.BFS_PUSH, labels=lab, timing=True)
        v = [l["advance_ms"] for l in st.levels]
        acc = v if acc is None else [x + y for x, y in zip(acc, v)]
    print([round(x / 5, 4) for x in acc])
    ctx.close()
