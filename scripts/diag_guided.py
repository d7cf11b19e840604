"""Diagnostic: where do the product and oracle guided steps diverge (per-parameter gradient error, block I/O gradients)."""
import sys, torch
sys.path.insert(0, ".")
import greedy_multimodal_learning_b200 as pkg
from oracle.mmtm_module import OracleMMTM
DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.deterministic = True
g = torch.Generator().manual_seed(224)
x = torch.randn(8, 2, 3, 224, 224, generator=g).to(DEV)
y = torch.randint(0, 40, (8,), generator=g).to(DEV)
res = {}
for name, cls in (("cuda", pkg.MMTM_mitigate), ("oracle", OracleMMTM), ("oracle2", OracleMMTM)):
    torch.manual_seed(777)
    model = pkg.MMTM_MVCNN(mmtm_cls=cls).to(DEV).train()
    io = {}
    for bn, blk in zip(("mmtm2", "mmtm3", "mmtm4"), model.mmtm_blocks()):
        def fh(mod, inp, out, bn=bn):
            io[bn + ".in0"] = inp[0].detach(); io[bn + ".out0"] = out[0].detach()
            inp[0].register_hook(lambda gr, bn=bn: io.__setitem__(bn + ".din0", gr.detach()))
            inp[1].register_hook(lambda gr, bn=bn: io.__setitem__(bn + ".din1", gr.detach()))
            out[0].register_hook(lambda gr, bn=bn: io.__setitem__(bn + ".dout0", gr.detach()))
            out[1].register_hook(lambda gr, bn=bn: io.__setitem__(bn + ".dout1", gr.detach()))
        blk.register_forward_hook(fh)
    fused, views, _, _ = model(x)
    loss = pkg.blend_loss(views, y)
    loss.backward()
    res[name] = (model, io)
for other in ("oracle2", "cuda"):
    print("== oracle vs", other)
    a, b = res[other], res["oracle"]
    for k in sorted(b[1]):
        sc = float(b[1][k].abs().max())
        print("  %-14s %.2e" % (k, float((a[1][k] - b[1][k]).abs().max()) / sc))
    pa, pb = dict(a[0].named_parameters()), dict(b[0].named_parameters())
    w = sorted(((float((pa[k].grad - pb[k].grad).abs().max()) / float(pb[k].grad.abs().max()), k) for k in pa), reverse=True)
    for e, k in w[:8]:
        print("  %.2e %s" % (e, k))
