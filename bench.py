#!/usr/bin/env python
"""bench.py — objects/s (N=1028 points) of the HS-Pose hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

One "step" = one train step over one synthetic batch per GPU: HSPose.forward
(augmentation, KNN + receptive-field gather + 3D-GCN backbone, dense heads,
fs_net + Chamfer losses) + backward + gradient all-reduce (N>1) + clip(5) + Adam.
Workload = BASELINE.json configs[2]: batch 128 per GPU, N=1028, k=20, S=7, bf16
autocast on the dense GEMMs, fp32 KNN / graph-conv kernels.  Weak scaling: the
per-GPU batch is fixed, objects are sharded over ranks, no data-path collective.

`value`  : objects/s with the batch already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same step through the public API with HOST (pinned) inputs — H2D of the
           12 input tensors and D2H of the loss inside the timed region.
`roofline`: the dominant hand-written kernel of the step, timed live with CUDA events on its
           launching stream; achieved = algorithmic bytes / launch time vs MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port (oracle/train_step.py) of the same step on the host cores,
           bounded sample.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "objects/sec (N=1028 pts) fwd+bwd"
UNIT = "objects/s"
N_PTS, K_NBR, S_SUP = 1028, 20, 7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"],
                    help="reference = the unmodified reference (oracle/_ref) on the host cores; reference-gpu = the "
                         "same module on cuda:0 (informational GPU denominator)")
    ap.add_argument("--optimizer", default="adam", choices=["adam", "ranger"],
                    help="adam (BASELINE.json configs[2]) or ranger (the reference's optimiser)")
    ap.add_argument("--batch", type=int, default=128, help="objects per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=4, help="objects per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of one CUDA graph")
    ap.add_argument("--losses", default="all", choices=["all", "fsnet"],
                    help="all = the 19 terms engine/train.py sums (fs_net + recon_6face + geo + prop) + Chamfer; "
                         "fsnet = fs_net + Chamfer only (the round-1 workload)")
    return ap.parse_args()


def workload_config(args, world):
    losses = "fs_net + recon_6face + geo + prop (19 terms) + Chamfer" if args.losses == "all" else "fs_net + Chamfer"
    return {"workload": f"train step (fwd+bwd+clip+{args.optimizer.capitalize()}, {losses} losses) batch={args.batch}/GPU "
                        f"N={N_PTS} k={K_NBR} S={S_SUP}",
            "global_batch": args.batch * world, "n_points": N_PTS, "k": K_NBR,
            "precision": "bf16 autocast dense GEMMs; fp32 KNN/graph-conv kernels"
            if args.precision == "bf16" else "fp32",
            "parallelism": f"dp{world} (batch sharded, one NCCL grad all-reduce)",
            "launch": "eager" if args.no_graph else "whole step replayed as one CUDA graph",
            "l2": "per-step working set (>2 GB of activations at batch 128) exceeds the 126 MB L2; "
                  "no explicit flush"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons (NVML) during the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- roofline
def _alg_bytes(name, a, train=True):
    """ALGORITHMIC bytes of one launch (each kernel-boundary input read once, each output
    written once; fp32, int32 indices) — the per-kernel split of SURVEY.md §8(d)'s formula."""
    if name == "hsp_graph_conv_fwd":
        dt, B, N, k, S, C = a[:6]
        e = 2 if dt == 1 else 4   # bytes per element of P
        return B * (12 * N + 4 * N * k + e * N * (S + 1) * C + 4 * N * C + (N * S * C if train else 0))
    if name == "hsp_graph_conv_bwd":
        dt, B, N, k, S, C = a[:6]
        e = 2 if dt == 1 else 4
        return B * (12 * N + 4 * N * k + e * N * (S + 1) * C + N * S * C + 4 * N * C + 4 * N * (S + 1) * C)
    if name == "hsp_graph_conv_bwd_obj":   # ints: p_dtype, B, N, k, S, C, gp_dtype, ws
        dt, B, N, k, S, C, gdt = a[:7]
        e, eo = (2 if dt == 1 else 4), (2 if gdt == 1 else 4)
        return B * (12 * N + 4 * N * k + e * N * (S + 1) * C + N * S * C + 4 * N * C + eo * N * (S + 1) * C)
    if name == "hsp_surface_conv_fwd" or name == "hsp_surface_conv_bwd":
        B, N, k, S, C = a[:5]
        return B * (12 * N + 4 * N * k + 4 * N * C)
    if name == "hsp_knn_feat":
        B, N, D, k = a[:4]
        return B * (4 * N * D + 4 * N * k)
    if name == "hsp_knn3":
        B, M, N, k = a[:4]
        return B * (12 * N + 12 * M + 4 * M * k)
    if name in ("hsp_gather_max_fwd", "hsp_gather_max_bwd"):
        B, N, C, R, kuse, kstride = a[:6]
        return B * (4 * N * C + 4 * R * kuse + 5 * R * C)
    if name in ("hsp_orl_global_fwd", "hsp_orl_global_bwd"):
        B, N, C, k = a[:4]
        return B * (4 * N * C + 4 * N * k + N * C + 4 * C)
    if name in ("hsp_upsample_rows_fwd", "hsp_upsample_rows_bwd"):
        B, Nsrc, M, C = a[:4]
        return B * (4 * Nsrc * C + 4 * M * C + 4 * M)
    # BatchNorm: COMPULSORY traffic only (x read once, y written once; backward x, dy read once, dx written) —
    # the kernels' own second read of x / dy is overhead, not algorithmic bytes
    if name == "hsp_bn_relu_fwd":      # ints: ldx, M, C, dtype, relu, ldy, ws
        ldx, M, C, dt = a[:4]
        return 2 * M * C * (2 if dt == 1 else 4)
    if name == "hsp_bn_apply_fwd":     # ints: ldx, M, C, dtype, nblocks, ldp, relu, ldy
        ldx, M, C, dt = a[:4]
        return 2 * M * C * (2 if dt == 1 else 4)
    if name == "hsp_bn_relu_bwd":      # ints: ldx, lddy, M, C, dtype, ...
        ldx, lddy, M, C, dt = a[:5]
        return 3 * M * C * (2 if dt == 1 else 4)
    if name == "hsp_gemm_bf16":        # ints: lda, a_mn, ldb, b_mn, M, N, K, rows_per_group, relu, ldo, out_f32, splits, ...
        lda, amn, ldb, bmn, M, N, K = a[:7]
        f32, splits = (a[10], a[11]) if len(a) > 11 else (0, 1)
        return 2 * (M * K + N * K) + (4 * splits if f32 else 2) * M * N
    if name in ("hsp_chamfer_fwd", "hsp_chamfer_bwd"):
        B, N, M = a[:3]
        return B * (12 * (N + M) + 8 * (N + M))
    return 0


def _alg_flops(name, a):
    """Arithmetic of one launch (multiply-add = 2 flops), for the competing compute bound."""
    if name == "hsp_gemm_bf16":
        lda, amn, ldb, bmn, M, N, K = a[:7]
        return 2.0 * M * N * K
    if name in ("hsp_graph_conv_fwd", "hsp_graph_conv_bwd", "hsp_graph_conv_bwd_obj"):
        dt, B, N, k, S, C = a[:6]
        per = 8.0 if name.endswith("fwd") else 8.0 / k      # theta (3 FMA) + product + max per (n,k,s,c); bwd: winners only
        return B * N * k * S * C * per
    if name in ("hsp_surface_conv_fwd", "hsp_surface_conv_bwd"):
        B, N, k, S, C = a[:5]
        return B * N * k * S * C * (7.0 if name.endswith("fwd") else 7.0 / k)
    if name == "hsp_knn_feat":
        B, N, D, k = a[:4]
        return B * (2.0 * N * N * D + 3.0 * N * N)
    if name == "hsp_knn3":
        B, M, N, k = a[:4]
        return B * (2.0 * M * N * 3 + 3.0 * M * N)
    return 0.0


def _gather_bytes(name, a):
    """SM <-> L2 bytes the ALGORITHM needs (every gathered row counted once per use): the gather kernels'
    real yardstick ('the gather is L2-bandwidth-, not HBM-bound')."""
    if name in ("hsp_graph_conv_fwd", "hsp_graph_conv_bwd", "hsp_graph_conv_bwd_obj"):
        dt, B, N, k, S, C = a[:6]
        e = 2 if dt == 1 else 4
        if name.endswith("fwd"):
            return B * N * k * S * C * e
        if name.endswith("obj"):                  # winners only: P read, arg-max byte, 16-byte pair record per 32 channels
            return B * N * S * C * (e + 1) + B * N * S * (C // 32) * k * 16
        return B * N * S * C * (e + 4 + 1)       # winners only: P read, gP atomic, arg-max byte
    if name in ("hsp_orl_global_fwd",):
        B, N, C, k = a[:4]
        return B * N * k * C * 4
    return 0


def kernel_breakdown(records, steps):
    agg = {}
    for name, a, ms in records:
        key = (name, a[:12])
        d = agg.setdefault(key, {"ms": 0.0, "n": 0})
        d["ms"] += ms
        d["n"] += 1
    rows = []
    for (name, a), d in agg.items():
        rows.append({"kernel": name, "dims": list(a), "launches_per_step": d["n"] / steps,
                     "ms_per_launch": d["ms"] / d["n"], "ms_per_step": d["ms"] / steps,
                     "alg_bytes": _alg_bytes(name, a), "alg_flops": _alg_flops(name, a),
                     "gather_bytes": _gather_bytes(name, a)})
    rows.sort(key=lambda r: -r["ms_per_step"])
    return rows


# entry point -> (ncu kernel-name prefix, grid) of its dominant kernel, for the DRAM-traffic lookup
def _ncu_key(name, a):
    if name == "hsp_graph_conv_bwd":
        dt, B, N, k, S, C = a[:6]
        return "graph_conv_bwd2_kernel", f"({min(B * ((N + 7) // 8), 592)}, 1, {(C + 127) // 128})"
    if name == "hsp_graph_conv_bwd_obj":
        dt, B, N, k, S, C = a[:6]
        return "graph_conv_bwd_obj_kernel", f"({C // 32}, {S + 1}, {B})"
    if name == "hsp_graph_conv_fwd":
        dt, B, N, k, S, C = a[:6]
        pairs = C // 2
        lanes = min(pairs, 128)
        return "graph_conv_fwd2_kernel", f"({(N + 7) // 8}, {B}, {(pairs + lanes - 1) // lanes})"
    if name == "hsp_knn_feat":
        B, N, D, k = a[:4]
        if D in (128, 256) and N >= 128:
            return "tc::knn_feat_tc2_kernel", f"({(N + 127) // 128}, {B}, 1)"
        return "knn_feat_kernel", f"({(N + 63) // 64}, {B}, 1)"
    if name == "hsp_surface_conv_fwd":
        B, N, k, S, C = a[:5]
        return "surface_conv_fwd2_kernel", f"({(N + 7) // 8}, {B}, 1)"
    if name == "hsp_knn3":
        B, M, N, k = a[:4]
        return "knn3_reg_kernel", None
    if name == "hsp_bn_relu_bwd":
        return "bn_bwd_apply_kernel", None
    if name == "hsp_bn_relu_fwd":
        return "bn_apply_kernel", None
    return None, None


def dram_traffic(name, a, ms_per_launch=None):
    """dram__bytes_read.sum + dram__bytes_write.sum of the entry point's dominant kernel, per launch,
    from the committed `ncu --set full` capture (profiles/kernel_traffic.json); None if not captured.
    The GEMM kernel serves every shape with the same grid: its launch is identified by duration (within 15 %)."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            table = json.load(f)
    except Exception:
        return None, None
    if name == "hsp_gemm_bf16" and ms_per_launch:
        best = None
        for e in table:
            if e["kernel"].startswith("gemm::gemm_tc_kernel") and e.get("duration_us"):
                d = abs(e["duration_us"] - ms_per_launch * 1e3) / (ms_per_launch * 1e3)
                if d <= 0.15 and (best is None or d < best[0]):
                    best = (d, e)
        return (best[1]["dram_bytes"], best[1].get("source")) if best else (None, None)
    prefix, grid = _ncu_key(name, a)
    if prefix is None:
        return None, None
    for e in table:
        if e["kernel"].startswith(prefix) and (grid is None or e["grid"] == grid):
            return e["dram_bytes"], e.get("source")
    return None, None


# what actually bounds each hand-written kernel (ncu evidence: profiles/*_ncu_full_summary.md)
LIMITER = {
    "hsp_gemm_bf16": "tcgen05 tensor pipe at the power-capped clock; TMA tensor loads L2 -> shared (5-stage ring), "
                     "epilogue (tcgen05.ld -> bf16 -> TMA store + BatchNorm partials) overlapped via 2 TMEM buffers",
    "hsp_graph_conv_bwd": "L2 atomic (RED.f32) throughput: N*S*C scattered adds per object; DRAM 16-19 %, issue 16-18 %",
    "hsp_graph_conv_bwd_obj": "winner byte -> pair record -> support value latency chain (35 % of stall samples) + shared-memory "
                              "atomics (23 %); slab written once, no memset / RED / cast pass",
    "hsp_graph_conv_fwd": "instruction issue (N*k*S*C element ops from L2-resident rows); DRAM 7 %, issue 68 %",
    "hsp_knn_feat": "L2 -> SM streaming of the candidate tiles (two sweeps) in the tcgen05 filter + exact-refine L2 row gathers",
    "hsp_knn3": "ALU pipe (selection network); DRAM 0 %",
    "hsp_surface_conv_fwd": "ALU/FMA issue; DRAM 3 %",
    "hsp_bn_relu_bwd": "HBM stream (5 passes over the activation matrix), 75 % of measured copy bandwidth at 1024 channels",
    "hsp_bn_relu_fwd": "HBM stream (3 passes), 63-66 % of measured copy bandwidth",
}


def peaks():
    """(hbm GB/s, bf16 TFLOP/s sustained, source) — MEASURED_PEAKS.json, else the recipe's stated fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md: 6.65 TB/s, ~1.4 PFLOP/s sustained)"


def measure_l2_gbs(dev):
    """SM <-> L2 copy bandwidth measured live (read + write of a working set that stays in the 126 MB L2):
    the yardstick of the gather kernels, which never reach HBM (no such figure in MEASURED_PEAKS.json)."""
    import torch
    a = torch.empty(6 * 1024 * 1024, dtype=torch.float32, device=dev)     # 24 MB + 24 MB
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    return 50 * 2 * a.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9


def bounds_of(row, hbm_gbs, bf16_tf, fp32_tf, l2_gbs):
    """Competing lower bounds of one launch (ms) and which one binds."""
    b = {"hbm_ms": row["alg_bytes"] / (hbm_gbs * 1e9) * 1e3}
    if row["kernel"] == "hsp_gemm_bf16":
        b["tensor_ms"] = row["alg_flops"] / (bf16_tf * 1e12) * 1e3
    elif row["alg_flops"]:
        b["fp32_ms"] = row["alg_flops"] / (fp32_tf * 1e12) * 1e3
    if row.get("gather_bytes"):
        b["l2_gather_ms"] = row["gather_bytes"] / (l2_gbs * 1e9) * 1e3
    binding = max(b, key=b.get)
    return {"bounds_ms": {k: round(v, 4) for k, v in b.items()}, "binding": binding.replace("_ms", ""),
            "frac_of_binding_bound": round(b[binding] / row["ms_per_launch"], 4)}


# ----------------------------------------------------------------------------- reference / CPU arm
def reference_throughput(sample_objects, steps, warmup, device="cpu", seed=1):
    """The train step of the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.sh), driven exactly
    as engine/train.py:76-110 drives it: HSPose('PoseNet_only').forward(do_loss=True), sum of the 19 loss
    terms, backward, clip_grad_norm_(5), Ranger.step — on the host cores (all threads) or, informationally,
    on cuda:0.  Falls back to the oracle port (oracle/train_step.py) when the staged tree is absent.
    -> (objects/s, seconds per step, kind, description)"""
    import warnings

    import torch
    from hspose_b200.synth import synth_batch
    torch.set_num_threads(os.cpu_count() or 1)
    batch = synth_batch(sample_objects, N_PTS, seed=seed, train=True)
    try:
        from oracle.ref_loader import import_reference
        FLAGS, ref_hspose, ranger2020, where = import_reference(train=1)
    except Exception as e:                       # staged tree absent: the restatement stands in
        if device != "cpu":
            raise SystemExit(f"reference-gpu needs the staged reference (oracle/make_ref.sh): {e}")
        from hspose_b200.HSPose import HSPose
        from oracle.train_step import OracleTrainer
        torch.manual_seed(0)
        tr = OracleTrainer(HSPose("PoseNet_only").state_dict())
        torch.manual_seed(1234)
        for _ in range(warmup):
            tr.step(batch)
        t0 = time.perf_counter()
        for _ in range(steps):
            tr.step(batch)
        dt = (time.perf_counter() - t0) / max(steps, 1)
        return sample_objects / dt, dt, "port", "oracle/train_step.py (materialising PyTorch port; fs_net + Chamfer losses, Adam)"
    warnings.filterwarnings("ignore")
    dev = torch.device(device)
    if dev.type == "cuda":
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    net = ref_hspose.HSPose("PoseNet_only").to(dev).train()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        opt = ranger2020.Ranger(net.build_params(training_stage_freeze=[])[0]["params"], lr=1e-4)
    batch = {k: v.to(dev) for k, v in batch.items()}

    def step():
        out, losses = net(**batch, do_loss=True)
        total = sum(v for grp in losses.values() for v in grp.values() if torch.is_tensor(v))
        opt.zero_grad()
        total.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 5)
        opt.step()
        return float(total)
    torch.manual_seed(1234)
    for _ in range(warmup):
        step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return sample_objects / dt, dt, "reference", (f"unmodified reference ({where}): network.HSPose.HSPose('PoseNet_only') "
                                                  f"fwd + 19 losses + bwd + clip_grad_norm_(5) + Ranger.step, fp32, on {device}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    gpu = args.impl == "reference-gpu"
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    n_obj = 32 if gpu else args.cpu_sample
    val, dt, kind, what = reference_throughput(n_obj, steps, warm, device="cuda:0" if gpu else "cpu")
    sample = (f"{steps} timed + {warm} warm-up train steps of {n_obj} objects (N={N_PTS}) — bounded sample of the "
              f"batch-{args.batch} workload; {what}")
    line = {"impl": args.impl, "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, max(world, 1)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count() if not gpu else 0, "kind": kind,
                             "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import hspose_b200.ops as ops
    from hspose_b200 import _lib, parallel
    from hspose_b200.engine import TrainStep
    from hspose_b200.HSPose import HSPose
    from hspose_b200.synth import synth_batch  # seeded synthetic inputs only (no oracle compute)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    rank, world, local = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _lib.check(_lib.load().hsp_device_check(), "hsp_device_check")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    torch.manual_seed(0)                       # identical weights on every rank
    groups = ("fsnet", "recon", "geo", "prop") if args.losses == "all" else ("fsnet",)
    model = HSPose("PoseNet_only", chamfer_w=1.0, loss_groups=groups).to(dev).train()
    parallel.seed_all(1234)                    # same Pool_layer permutations on every rank
    amp = args.precision == "bf16"
    trainer = TrainStep(model, lr=1e-4, clip=5.0, amp=amp, graph=not args.no_graph, optimizer=args.optimizer)

    B = args.batch
    host = {k: v.pin_memory() for k, v in synth_batch(B, N_PTS, seed=1 + rank, train=True).items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def step(batch):
        return trainer(batch)

    def step_e2e():
        return trainer(host).item()            # H2D of the 12 inputs + D2H of the loss (4 bytes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    for _ in range(max(args.warmup, 3)):
        step(resident)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.launch_count()
    torch.cuda.profiler.start()    # `ncu --profile-from-start off` sees exactly the timed steps
    ms_step = timed(lambda: step(resident), args.steps)
    torch.cuda.profiler.stop()
    launches = ops.launch_count() - l0
    if trainer.launches_per_step is not None:   # graph replay: kernels recorded at capture time
        launches = trainer.launches_per_step * args.steps
    clocks = sampler.stop()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # per-kernel CUDA-event timing of the same step (separate instrumented pass)
    ksteps = min(3, args.steps)
    trainer.use_graph = False
    ops.enable_timing(True)
    for _ in range(ksteps):
        step(resident)
    torch.cuda.synchronize()
    rows = kernel_breakdown(ops.timing_records(), ksteps)
    ops.enable_timing(False)
    if os.environ.get("HSP_BENCH_DUMP") and rank == 0:     # every entry point x shape of the step, for analysis
        with open(os.environ["HSP_BENCH_DUMP"], "w") as f:
            json.dump(rows, f, indent=0)

    def leave():
        """Tear down: the CUDA graph that captured the NCCL all-reduces is destroyed BEFORE the process group
        (a communicator still referenced by a live graph blocks destroy_process_group)."""
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            trainer.graph = None
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.destroy_process_group()

    if rank != 0:
        leave()
        return

    peak, peak_tf, peak_src = peaks()
    l2_gbs = measure_l2_gbs(dev)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_tf = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12      # 148 SMs x 128 FMA lanes x 2 flops at the SAMPLED clock
    top = rows[0]
    own_ms = sum(r["ms_per_step"] for r in rows)
    traffic, traffic_src = dram_traffic(top["kernel"], tuple(top["dims"]), top.get("ms_per_launch"))
    if top["kernel"] == "hsp_gemm_bf16":     # a tensor-core kernel: its yardstick is the measured bf16 GEMM rate
        ach = top["alg_flops"] / (top["ms_per_launch"] * 1e-3) / 1e12
        bound, unit, pk = "tensor", "TFLOP/s", peak_tf
    else:
        ach = top["alg_bytes"] / (top["ms_per_launch"] * 1e-3) / 1e9
        bound, unit, pk = "hbm", "GB/s", peak
    roofline = {"bound": bound, "kernel": top["kernel"], "dims": top["dims"], "achieved": ach,
                "peak": pk, "unit": unit, "frac": ach / pk, "traffic": traffic,
                "traffic_source": traffic_src, "limiter": LIMITER.get(top["kernel"]),
                "peak_source": peak_src, "alg_bytes_per_launch": top["alg_bytes"],
                "alg_flops_per_launch": top["alg_flops"],
                "ms_per_launch": top["ms_per_launch"], "share_of_step": top["ms_per_step"] / ms_step,
                "own_kernels_ms_per_step": own_ms,
                "yardsticks": {"hbm_gbs": peak, "bf16_tflops_sustained": peak_tf, "fp32_tflops_at_sampled_clock": fp32_tf,
                               "l2_copy_gbs_measured_live": l2_gbs},
                "top_kernels": [dict({"kernel": r["kernel"], "dims": r["dims"][:8],
                                      "ms_per_step": round(r["ms_per_step"], 4),
                                      "ms_per_launch": round(r["ms_per_launch"], 4),
                                      "hbm_GBps": round(r["alg_bytes"] / (r["ms_per_launch"] * 1e-3) / 1e9, 1),
                                      "TFLOPs": round(r["alg_flops"] / (r["ms_per_launch"] * 1e-3) / 1e12, 2)},
                                     **bounds_of(r, peak, peak_tf, fp32_tf, l2_gbs))
                                for r in rows[:10]]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        val, dt, kind, what = reference_throughput(args.cpu_sample, 3, 1)
        cpu = {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
               "sample": f"3 timed + 1 warm-up train steps of {args.cpu_sample} objects (N={N_PTS}), {what}, "
                         f"all host threads"}

    line = {"metric": METRIC, "value": B * world / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if amp else "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    leave()


if __name__ == "__main__":
    a = parse()
    if a.impl in ("reference", "reference-gpu"):
        run_reference(a)
    else:
        run_b200(a)
