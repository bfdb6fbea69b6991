#!/usr/bin/env python
"""Headline benchmark: crystals/sec sampled with the full 1000-step reverse process (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one full pass of the hot path over one batch: the complete 1000-step predictor-corrector
reverse diffusion (2 000 score-network evaluations) of `--batch` crystals per GPU, DiffCSP CSPNet at the
upstream default size (hidden 512, 6 layers, 128 frequencies, fully-connected edges), synthetic
random-init weights and mp_20 atom counts.  One JSON line is printed by rank 0.

value      device-resident throughput, WEAK scaling: `--batch` (256, the batch size of BASELINE configs[1]) crystals per
           GPU, the N x 256 drawn crystals cut into contiguous shards balanced by sum n^2 (CUDA events, max over ranks;
           Philox noise generated in-graph)
strong_scaling   the same sampler on a FIXED global batch of 1024 crystals (BASELINE configs[2]) sharded over the N GPUs
           (N = 1 runs all 1024): the north-star's strong-scaling axis
e2e        same metric through the reference-facing plugin call DiffCSPSampler.generate(): initial noise
           drawn on the HOST and copied H2D, per-step noise from torch's device generator, results
           post-processed and copied D2H — all inside the timed region
fine_tune_step   the reward-weighted fine-tune step (pipeline/mat_invent.py:125-189) at the reference's working point
           (18 crystals, accum_steps 50) sharded over the N GPUs, gradient all-reduce + flat Adam inside the timed epoch;
           at N > 1 also the all-reduce of the 49 MB gradient buffer on its own (us, bus GB/s)
roofline   dominant kernel (the per-edge GEMMs) — achieved algorithmic TFLOP/s over the measured peak
cpu_baseline  the UNMODIFIED reference sampler (models/diffcsp/diffusion.py DiffCSPModule.sample, staged under
           baseline/_ref/ by __graft_entry__.build(), imported under oracle/shims) on this box's host cores, bounded sample
gpu_reference   Comparator B: the same unmodified reference code `.cuda()` on this B200 — the same
           256-crystal batch, all 1000 steps — i.e. what a user of the reference gets on this box today

`--impl reference` times only the reference's CPU path and prints the same line shape.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HP = dict(hidden_dim=512, num_layers=6, num_freqs=128, max_atoms=100, time_dim=256, latent_dim=0,
          timesteps=1000, sigma_begin=0.005, sigma_end=0.5, costs=(1.0, 1.0, 20.0))
STEP_LR = 5e-6
HEAD_SCALE = 0.05          # output heads x0.05 keep random-init trajectories finite


def atom_counts(total):
    from matinvent_b200.models.diffcsp.sample import ATOM_DIST
    return np.random.RandomState(0).choice(21, total, p=ATOM_DIST["mp_20"]).tolist()


def sigmas_norm():
    p = os.path.join(ROOT, "tests", "golden", "sigmas_norm_T1000.pt")
    return torch.load(p)["sigmas_norm"]


def build_model(device):
    from matinvent_b200.models.diffcsp import DiffCSPModule
    m = DiffCSPModule(
        decoder=dict(hidden_dim=HP["hidden_dim"], num_layers=HP["num_layers"], max_atoms=HP["max_atoms"],
                     num_freqs=HP["num_freqs"], edge_style="fc", cutoff=7.0, max_neighbors=20, ln=True, ip=True),
        beta_scheduler=dict(timesteps=HP["timesteps"], scheduler_mode="cosine"),
        sigma_scheduler=dict(timesteps=HP["timesteps"], sigma_begin=HP["sigma_begin"], sigma_end=HP["sigma_end"]),
        cost_lattice=HP["costs"][0], cost_coord=HP["costs"][1], cost_type=HP["costs"][2],
        time_dim=HP["time_dim"], latent_dim=HP["latent_dim"], device=device, sigmas_norm=sigmas_norm())
    m.decoder.reset_parameters(seed=0)
    for k in ("coord_w", "lattice_w", "type_w", "type_b"):
        m.decoder.w(k).mul_(HEAD_SCALE)
    m.decoder.weights_changed()
    return m


def flops_per_forward(na, full_heads=True):
    """Algorithmic FLOPs of one score-network evaluation as THIS implementation computes it
    (split first edge linear)."""
    H, F6, A, T, L = HP["hidden_dim"], 6 * HP["num_freqs"], HP["max_atoms"], HP["time_dim"], HP["num_layers"]
    N, E, B = sum(na), sum(n * n for n in na), len(na)
    edge = 2 * E * (F6 * H + H * H)
    node = 2 * N * (H * 2 * H + 2 * H * H + H * H) + 2 * B * 9 * H
    emb = 2 * N * (A * H + H * H) + 2 * B * T * H
    heads = 2 * N * 3 * H + (2 * N * A * H + 2 * B * 9 * H if full_heads else 0)
    return L * (edge + node) + emb + heads, L * edge


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 8]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm = [float(r[1]) for r in rows]
        busy = [v for v, r in zip(sm, rows) if float(r[3]) > 300.0] or sm
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k].lower().startswith("active") for r in rows)]
        return dict(sm_mhz=statistics.median(busy), sm_max_mhz=float(rows[0][2]), reasons=reasons,
                    power_w_max=max(float(r[3]) for r in rows), samples=len(rows))


# dram__bytes_read.sum + dram__bytes_write.sum of edge_pair_kernel<0> at the default workload, from the committed
# ncu --set full capture of the final build (profiles/r2x_b256_ncu_raw.csv; r2_edge_pair_ncu_raw.csv had 118.8 + 28.1 MB)
PAIR_TRAFFIC = 149.4e6      # 118.9 MB read + 30.5 MB written

WORKLOAD = "DiffCSP CSPNet(H512,L6,F128,fc) 1000-step sampler, batch=%d mp_20 crystals/GPU"


def reference_module(timesteps, device="cpu"):
    """the UNMODIFIED reference DiffCSPModule (full-size net, same seeded weights as build_model) from /root/reference or
    the staged baseline/_ref copy; None when neither exists"""
    from oracle import diffcsp_oracle as O
    from oracle import ref_import as R
    if not R.reference_available():
        return None
    hp = O.default_hparams(timesteps=timesteps)
    torch.manual_seed(1234)
    ref = R.build_reference_module(hp, sigmas_norm() if timesteps == HP["timesteps"] else None)
    ref.decoder.load_state_dict(O.init_params(hp, seed=0, head_scale=HEAD_SCALE))
    return ref.to(device).eval()


def cpu_reference_leg(na_all, state_dict=None, seconds_target=20.0, ncryst=64, steps=None):
    """The reference's CPU path on a bounded sample of the workload, all host threads: the first `ncryst` crystals,
    `steps` reverse steps (a complete DiffCSPModule.sample of a `steps`-step schedule: every reverse step costs the same —
    2 score-network evaluations + the update — so crystals/s for 1000 steps is steps-for-steps proportional).
    kind = "reference": the unmodified reference module; "port": the oracle restatement (only where no reference copy exists)."""
    from oracle import diffcsp_oracle as O
    from oracle.ref_import import make_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    na = [max(1, n) for n in na_all[:ncryst]]
    probe = reference_module(2)
    kind = "reference" if probe is not None else "port"

    def run(T):
        if kind == "reference":
            ref = probe if T == 2 else reference_module(T)
            t0 = time.perf_counter()
            with torch.no_grad():
                ref.sample(make_batch(na), step_lr=STEP_LR)
            return time.perf_counter() - t0
        hp = O.default_hparams()
        sd = state_dict if state_dict is not None else O.init_params(hp, seed=0, head_scale=HEAD_SCALE)
        sd = {k: v.detach().cpu().float() for k, v in sd.items()}
        t0 = time.perf_counter()
        with torch.no_grad():
            O.sample(sd, hp, O.Schedules(hp, sigmas_norm()), na, O.Noise(torch.Generator().manual_seed(0)), step_lr=STEP_LR, timesteps=T)
        return time.perf_counter() - t0

    per_step = run(2) / 2
    if steps is None:
        steps = int(max(4, min(200, seconds_target / max(per_step, 1e-3))))
    dt = run(steps)
    value = len(na) / (dt / steps * HP["timesteps"])
    return dict(value=value, unit="crystals/s", cores=cores, kind=kind,
                sample="%s DiffCSPModule.sample on the first %d crystals of the workload, %d reverse steps in %.1f s "
                       "(per-step cost is t-independent: x%.0f to 1000 steps)"
                       % ("unmodified reference" if kind == "reference" else "oracle port of", len(na), steps, dt,
                          HP["timesteps"] / steps)), dt / steps


def gpu_reference_leg(na, dev):
    """Comparator B: the unmodified reference sampler on this GPU, the benchmark batch, all 1000 steps, timed once."""
    from oracle.ref_import import make_batch
    ref = reference_module(HP["timesteps"], dev)
    if ref is None:
        return dict(unavailable="no reference copy (run __graft_entry__.build() where /root/reference exists)")
    batch = make_batch([max(1, n) for n in na])
    batch.num_atoms, batch.batch = batch.num_atoms.to(dev), batch.batch.to(dev)
    with torch.no_grad():
        warm = reference_module(4, dev)
        warm.sample(batch, step_lr=STEP_LR)                    # cuBLAS / allocator warm-up on a 4-step schedule
        del warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, _ = ref.sample(batch, step_lr=STEP_LR)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    ok = bool(torch.isfinite(out["lattices"]).all())
    del ref, out
    torch.cuda.empty_cache()
    return dict(value=len(na) / dt, unit="crystals/s", seconds=dt, crystals=len(na), reverse_steps=HP["timesteps"], finite=ok,
                note="unmodified reference models/diffcsp/{diffusion,cspnet}.py on cuda (torch %s eager, fp32, TF32 off as "
                     "at the reference's first sample_step), full 1000 steps, one pass" % torch.__version__)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    na = atom_counts(args.batch)
    vals, ms = [], []
    total = args.warmup + args.steps
    for i in range(total):
        cb, per_step = cpu_reference_leg(na, seconds_target=max(3.0, 150.0 / total))
        if i >= args.warmup:
            vals.append(cb["value"])
            ms.append(per_step * 1e3)
    v = float(np.mean(vals))
    cb["value"] = v
    print(json.dumps(dict(
        metric="crystals/sec sampled (1000-step reverse)", value=v, unit="crystals/s", impl="reference",
        n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=float(np.mean(ms)) * HP["timesteps"],
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=WORKLOAD % args.batch,
                    note="the reference's CPU implementation of the path on this box's host cores; every step is a bounded "
                         "sample of the workload (cpu_baseline.sample)"),
        cpu_baseline=cb, e2e=dict(value=v, unit="crystals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="crystals per GPU (the batch size of BASELINE configs[1])")
    ap.add_argument("--strong-batch", type=int, default=1024, help="global batch of the strong-scaling leg (BASELINE configs[2])")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-gpu-ref", action="store_true", help="skip Comparator B (the unmodified reference on this GPU)")
    ap.add_argument("--timesteps", type=int, default=None, help="debug: shorter reverse process (invalid as a bench value)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ffma", action="store_true", help="disable the tensor-core GEMMs (FP32 CUDA-core path)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries the ONE JSON line and nothing else: NCCL writes its version banner to file descriptor 1 when the
    # communicator comes up (NCCL_DEBUG_FILE does not catch it on every box), so fd 1 points at stderr for the whole
    # run and the result goes out through the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from matinvent_b200 import _lib
    from matinvent_b200.models.diffcsp import DiffCSPSampler, PhiloxNoise
    from matinvent_b200.models.diffcsp.sample import CrystalBatch, CrystalData
    from matinvent_b200 import ops

    m = build_model(dev)
    m.decoder.use_tc = not args.ffma
    T = args.timesteps or HP["timesteps"]
    from matinvent_b200.models.diffcsp.finetune import partition_crystals
    na_all = atom_counts(args.batch * world)
    # weak scaling: `--batch` crystals per GPU on average; the shards are contiguous and balanced by sum n^2 (edges),
    # so the max-over-ranks time measures the sampler, not the luck of the draw (a cut by count leaves +-12 % edges)
    lo_, hi_ = partition_crystals(na_all, world)[rank]
    na = na_all[lo_:hi_]
    batch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(seed):
        out, _ = m.sample(batch, step_lr=STEP_LR, noise=PhiloxNoise(dev, seed=seed), timesteps=args.timesteps)
        return out

    for w in range(args.warmup):
        one_pass(100 + w)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    l0 = _lib.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for k in range(args.steps):
        out = one_pass(k)
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1])
    clk = clocks.stop()
    # kernels in the timed region: eager first step + one graph capture's worth per replayed step
    g = m.decoder.graph_for(batch.num_atoms)
    per_step_launches = None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    value = len(na_all) * args.steps / (ms / 1e3) * (HP["timesteps"] / T)
    assert torch.isfinite(out["lattices"]).all() and torch.isfinite(out["frac_coords"]).all()

    # ---- launches: count one eager reverse step, multiply (graph replays launch the same kernels)
    from matinvent_b200.models.diffcsp.diffusion import _StepState
    from matinvent_b200.models.diffcsp.scheduler import StepCoefficients
    co = StepCoefficients(m.beta_scheduler, m.sigma_scheduler, STEP_LR)
    st = _StepState(m, g, out["frac_coords"].clone(), out["lattices"].clone(), out["atom_types"].clone(), co, T,
                    PhiloxNoise(dev, seed=1), False)
    l1 = _lib.launches
    st.step(last=False)
    per_step_launches = _lib.launches - l1
    gpu_launches = per_step_launches * T * args.steps

    # ---- roofline of the dominant kernel: the per-edge GEMMs, CUDA events around each launch of one
    # eager score-network evaluation on the launching stream
    flops_fwd, flops_edge = flops_per_forward(na)
    H, F6 = HP["hidden_dim"], 6 * HP["num_freqs"]
    ws = m.decoder.workspace(g, False)
    W = m.decoder.w
    dec = m.decoder
    presplit, merged = dec.edge_mode(g.E)
    torch.cuda.synchronize()
    evs = []
    for rep in range(3):
        for i in range(HP["num_layers"]):
            q = "l%d." % i
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            if merged:
                ws.cat[0][:, H:].zero_()          # destination of the fused scatter-mean (the layer's LayerNorm zeroes it in situ)
            e0.record()
            dec.edge_gemm1(i, ws, g, g.E, ws.a1[0], False, presplit, merged)
            e1.record()
            dec.edge_gemm2(i, ws, g, g.E, ws.a1[0], ws.cat[0][:, H:], False, merged)
            e2.record()
            evs.append((e0, e1, e2))
    torch.cuda.synchronize()
    t_pair = statistics.mean(a.elapsed_time(c) for a, b, c in evs[HP["num_layers"]:]) / 1e3     # s per (GEMM1+GEMM2)
    t_g1 = statistics.mean(a.elapsed_time(b) for a, b, c in evs[HP["num_layers"]:]) * 1e3      # us
    t_g2 = statistics.mean(b.elapsed_time(c) for a, b, c in evs[HP["num_layers"]:]) * 1e3
    fl_pair = 2 * g.E * (F6 * H + H * H)
    fl_g1 = 2 * g.E * F6 * H
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    ach = fl_g1 / (t_g1 * 1e-6) / 1e12
    ach_pair = fl_pair / t_pair / 1e12
    pair = dec.pair_mode(False, merged)
    kname = ("edge_pair_kernel<0> (tcgen05.mma.cta_group::2, split-precision FP16 x3, both operands staged by TMA)" if pair else
             "tc_gemm_kernel (tcgen05, split-precision FP16 x3)" if dec.use_tc else "sgemm_kernel (FP32 FFMA)")
    # DRAM traffic of this launch from the committed ncu --set full capture of the default workload
    # (profiles/r2x_b256_ncu_raw.csv, edge_pair_kernel<0>)
    traffic = PAIR_TRAFFIC if (pair and g.E == 34445) else None
    roofline = dict(bound="tensor", kernel=kname + ": per-edge GEMM 1 (Phi.W_F^T + 2 gathered rows + SiLU), the largest launch",
                    achieved=ach, peak=peak_tf, unit="TFLOP/s", frac=ach / peak_tf, traffic=traffic,
                    flop_per_launch=fl_g1, us_per_launch=t_g1,
                    algorithmic_bytes_per_launch=2 * 2 * g.E * F6 + 4 * g.E * H + 2 * 2 * H * F6 + 2 * 4 * g.N * 2 * H,
                    peak_source="MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                    note="achieved = algorithmic FP32 FLOPs / CUDA-event time of the launch; FP32-grade accuracy costs 3 fp16 MMAs per "
                         "product (x = hi + lo), so the ceiling of frac is 1/3 (1e-4 parity over 2000 chained forwards rules out "
                         "plain TF32/BF16/FP16 inputs); tiles: %s; edge GEMM pair (this launch + the K=512 one): %.1f TFLOP/s, "
                         "share of step = %.2f"
                         % ("256x256 per CTA pair, one accumulator" if pair else "128x256, one accumulator" if merged else
                            "128x128, main + correction accumulators", ach_pair,
                            2 * HP["num_layers"] * t_pair / (ms / 1e3 / args.steps / T)),
                    mma_tflops=3 * ach, frac_mma_of_peak=3 * ach / peak_tf, us_gemm1=t_g1, us_gemm2=t_g2)
    # the edge-scatter (segment-mean) kernel against the HBM roofline.  Two timings: (a) back to back over rotating inputs that
    # together exceed the L2 several times (6 x E x H floats), i.e. as the kernel runs inside a step — launch latency and ramp
    # overlap the previous launch; (b) one launch at a time after a 256 MB memset (whose dirty lines are written back while the
    # kernel reads, and whose launch ramp is fully exposed)
    if ws.a2.shape[0] < g.E:
        ws.a2 = torch.empty(g.E, H, device=dev)
    xs_rot = [ws.a2] + [torch.randn(g.E, H, device=dev) for _ in range(5)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    seg = []
    for rep in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.segment_reduce(ws.a2, g.seg_ptr, ws.cat[0][:, H:], g.N, H, mean=True)
        b.record()
        seg.append((a, b))
    torch.cuda.synchronize()
    t_iso = statistics.median(a.elapsed_time(b) for a, b in seg[1:]) / 1e3
    reps = 10
    for _ in range(2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for r_ in range(reps):
            for X_ in xs_rot:
                ops.segment_reduce(X_, g.seg_ptr, ws.cat[0][:, H:], g.N, H, mean=True)
        b.record()
        torch.cuda.synchronize()
    t_seg = a.elapsed_time(b) / 1e3 / (reps * len(xs_rot))
    del xs_rot
    seg_bytes = 4 * g.E * H + 4 * (g.N + 1) + 4 * g.N * H
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline_scatter = dict(bound="hbm", kernel="segment_reduce_kernel (edge scatter-mean, standalone: knn graphs and the backward; the "
                                                "inference path forms the scatter-mean in the second per-edge GEMM's epilogue)",
                            achieved=seg_bytes / t_seg / 1e9, peak=hbm_peak, unit="GB/s", frac=seg_bytes / t_seg / 1e9 / hbm_peak,
                            traffic=71.8e6 if g.E == 34445 else None,      # ncu: profiles/r1_final_segment_reduce_ncu_raw.csv
                            bytes_per_launch=seg_bytes, us_per_launch=t_seg * 1e6,
                            isolated=dict(us_per_launch=t_iso * 1e6, achieved=seg_bytes / t_iso / 1e9, frac=seg_bytes / t_iso / 1e9 / hbm_peak),
                            note="achieved: 60 launches back to back over 6 rotating inputs (6 x %d MB, far beyond the 126 MB L2), CUDA "
                                 "events around the sequence; isolated: one launch after a 256 MB memset (dirty-line write-back and the "
                                 "launch ramp inside the timed window: ~10 us of fixed cost on a 12 us transfer)" % (4 * g.E * H >> 20))

    # ---- strong scaling: a FIXED global batch (BASELINE configs[2]: 1024 crystals) sharded over the N GPUs
    strong = None
    if not args.no_strong and not args.timesteps:
        na_s = atom_counts(args.strong_batch)
        parts = partition_crystals(na_s, world)
        slo, shi = parts[rank]
        sbatch = CrystalBatch([CrystalData(None, None, None, None, n) for n in na_s[slo:shi]])
        m.sample(sbatch, step_lr=STEP_LR, noise=PhiloxNoise(dev, seed=50))            # warm: workspace + graph capture
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_s = 2
        a.record()
        for k in range(n_s):
            m.sample(sbatch, step_lr=STEP_LR, noise=PhiloxNoise(dev, seed=60 + k))
        b.record()
        barrier()
        ts = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e_rank = [sum(n * n for n in na_s[p:q]) for p, q in parts]
        strong = dict(value=args.strong_batch * n_s / (float(ts) / 1e3), unit="crystals/s", global_batch=args.strong_batch,
                      crystals_per_gpu=[q - p for p, q in parts], edges_per_gpu=e_rank, passes=n_s,
                      ms_per_pass=float(ts) / n_s, scaling="strong",
                      note="BASELINE configs[2]: fixed global batch sharded by sum n^2, no collective while sampling; "
                           "speed-up over N = 1 is this value / the N = 1 line's strong_scaling.value")
        m.decoder.release(m.decoder.graph_for(sbatch.num_atoms))

    # ---- e2e through the plugin call, host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        sampler = DiffCSPSampler(batch_size=len(na), num_batches=1)
        torch.manual_seed(1234 + rank)
        gen_kw = {}

        def same_workload():
            # generate() draws its atom counts from numpy's global RNG (sample.py:117-138).  One GPU: seed the stream so
            # that it draws exactly the benchmark workload (atom_counts); several GPUs: hand every rank its shard of that
            # draw, which is what MatInvent.sample_step does
            np.random.seed(0)
            if world > 1:
                gen_kw["num_atoms"] = na

        if args.timesteps:
            e2e = None
        else:
            same_workload()
            sampler.generate(m, **gen_kw)               # warm
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            n_e2e = max(1, min(args.steps, 2))
            for _ in range(n_e2e):
                same_workload()
                data, _ = sampler.generate(m, **gen_kw)
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            nn = sum(int(d.num_atoms) for d in data)
            e2e = dict(value=len(na_all) * n_e2e / (float(t) / 1e3), unit="crystals/s",
                       h2d_bytes_per_step=4 * (nn * 3 + len(na) * 9 + nn * 100),
                       d2h_bytes_per_step=4 * (nn * 3 + nn + len(na) * 6), passes=n_e2e,
                       same_workload_as_value=bool(nn == g.N))

    # ---- second half of the hot path: the reward-weighted fine-tune step at the reference's working point
    # (<= 18 crystals, accum_steps 50), CUDA events around one epoch of 1000 timesteps as the pipeline runs
    # it: the first group of stacked timesteps eagerly, one CUDA-graph capture, 18 replays, 20 Adam steps
    ft = None
    if not args.no_e2e and not args.timesteps:
        from matinvent_b200.models.diffcsp.finetune import FineTuner
        prior = build_model(dev)
        for p_ in prior.parameters():
            p_.requires_grad = False
        gen = torch.Generator().manual_seed(7)
        nft = [max(1, n) for n in atom_counts(18)]
        crystals = []
        for n in nft:
            d_ = CrystalData(torch.rand(n, 3, generator=gen), torch.randint(1, 101, (n,), generator=gen),
                             3 + 5 * torch.rand(1, 3, generator=gen), 70 + 40 * torch.rand(1, 3, generator=gen), torch.tensor(n))
            d_.reward = torch.rand(1, generator=gen)
            crystals.append(d_)
        fbatch = CrystalBatch(crystals)
        snap = m.decoder.flat.data.clone()
        try:      # a secondary figure: it must never cost the headline line
            # every rank: same batch, same Philox seed (the noise is drawn for the GLOBAL batch and sliced), its own shard
            tuner = FineTuner(m, prior, lr=1e-4, accum_steps=50, sigma=0.025, noise=PhiloxNoise(dev, seed=3), rank=rank, world=world)
            tuner.run_batch(fbatch, 100)                       # warm: allocations, kernel attributes
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            tuner.run_batch(fbatch, 1000)
            b.record()
            barrier()
            tf = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            parts = partition_crystals(nft, world)
            ft = dict(ms_per_timestep=float(tf) / 1000, crystals=len(nft), atoms=sum(nft), edges=sum(n * n for n in nft),
                      crystals_per_gpu=[q - p for p, q in parts],
                      timesteps_per_launch=tuner.group_size(max(sum(n * n for n in nft[p:q]) for p, q in parts)), accum_steps=50,
                      adam_steps_in_epoch=20, grad_bytes=4 * tuner.grad.numel(),
                      note="agent forward + prior forward + losses + backward per timestep, every 50 timesteps ONE all-reduce of "
                           "the flat gradient buffer (N > 1) + the flat Adam kernel; one epoch of 1000 timesteps including its "
                           "CUDA-graph capture; max over ranks; the reference runs 3 such epochs per RL iteration")
            if world > 1:      # the collective on its own: the 49 MB flat gradient buffer over NVLink
                gbuf = tuner.grad
                for _ in range(3):
                    dist.all_reduce(gbuf)
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20):
                    dist.all_reduce(gbuf)
                b.record()
                barrier()
                ta = torch.tensor([a.elapsed_time(b) / 20], device=dev, dtype=torch.float64)
                dist.all_reduce(ta, op=dist.ReduceOp.MAX)
                nb = 4 * gbuf.numel()
                ft["allreduce_us"] = float(ta) * 1e3
                ft["allreduce_bus_gbs"] = 2 * (world - 1) / world * nb / (float(ta) / 1e3) / 1e9
                ft["allreduce_share_of_epoch"] = 20 * float(ta) / float(tf)
                gbuf.zero_()
        except Exception as exc:      # noqa: BLE001
            ft = dict(error="%s: %s" % (type(exc).__name__, exc))
        finally:
            m.decoder.flat.data.copy_(snap)                    # the benchmark model is left as it was
            m.decoder.weights_changed()
        del prior

    gpu_ref = None
    if rank == 0 and world == 1 and not args.no_gpu_ref and not args.timesteps and not args.no_cpu:
        try:
            gpu_ref = gpu_reference_leg(na_all, dev)
            if "value" in gpu_ref:
                gpu_ref["speedup_value_over_gpu_reference"] = value / gpu_ref["value"]
        except Exception as exc:      # noqa: BLE001
            gpu_ref = dict(error="%s: %s" % (type(exc).__name__, exc))

    cb = None
    if rank == 0 and not args.no_cpu:
        cb, _ = cpu_reference_leg(na_all, m.decoder.state_dict())

    if rank == 0:
        line = (json.dumps(dict(
            metric="crystals/sec sampled (1000-step reverse)", value=value, unit="crystals/s", n_gpus=world,
            steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=WORKLOAD % args.batch,
                        note="SURVEY.md §8(d) config (2): the DiffCSP back-end at the batch size of BASELINE configs[1]; the MatterGen "
                             "configs are unmeasured (its score network lives in an un-vendored package)",
                        crystals_per_gpu=len(na), atoms_per_gpu=g.N, edges_per_gpu=g.E, reverse_steps=T,
                        forwards_per_step=2, gflop_per_forward=flops_fwd / 1e9,
                        l2="per-step working set (weights 49 MB + Phi %d MB + 2x edge activations %d MB) exceeds the 126 MB L2; no flush"
                           % (4 * g.E * F6 >> 20, 2 * 4 * g.E * H >> 20),
                        parallelism="dp%d (crystals sharded by sum n^2, no collective while sampling; one gradient all-reduce per "
                                    "Adam step while fine-tuning)" % world),
            clocks=clk, gpu_launches=gpu_launches, launches_per_reverse_step=per_step_launches,
            e2e=e2e, strong_scaling=strong, fine_tune_step=ft, roofline=roofline, roofline_edge_scatter=roofline_scatter,
            cpu_baseline=cb, gpu_reference=gpu_ref)))
        sys.stdout.flush()
        os.write(json_fd, (line + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
