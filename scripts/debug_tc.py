import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData
dev = torch.device("cuda")
m = bench.build_model(dev)
dec = m.decoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1e4
na = bench.atom_counts(B)
g = dec.graph_for(na)
gen = torch.Generator().manual_seed(0)
temb = m.time_table()[500].expand(g.B, -1).contiguous()
a = (torch.randn(g.N, 100, generator=gen) * scale).cuda()
x = torch.rand(g.N, 3, generator=gen).cuda()
l = (torch.randn(g.B, 3, 3, generator=gen) * scale).cuda()
outs = {}
for mode in ("ffma", "tc"):
    dec.use_tc = mode == "tc"
    ws = dec.workspace(g, False)
    pl, px, pa = dec.forward_graph(g, temb, a, x, l)
    torch.cuda.synchronize()
    outs[mode] = dict(pl=pl.clone(), px=px.clone(), pa=pa.clone(), h0=ws.h0.clone(), h=ws.h[0].clone(), a1=ws.a1[0].clone(), a2=ws.a2.clone(),
                      cat=ws.cat[0].clone(), an1=ws.an1[0].clone(), pq=ws.pq.clone(), hf=ws.hf.clone(), amax=ws.amax.clone())
for k in outs["tc"]:
    t, f = outs["tc"][k], outs["ffma"][k]
    bad = (~torch.isfinite(t)).sum().item()
    err = float((t.double() - f.double()).abs().max() / f.double().abs().max().clamp_min(1e-30)) if bad == 0 else float("nan")
    print("%-5s max|ffma| %.3e  nonfinite(tc) %d  rel err %.2e" % (k, float(f.abs().max()), bad, err))
