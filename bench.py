#!/usr/bin/env python
"""Benchmark of the gated-fusion caption decoder hot path on B200 (BASELINE.json metric:
captions/sec, train fwd+bwd and greedy decode, MSRVTT-shape synthetic batches).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
    python bench.py --impl reference ...                     # the reference algorithm on the host CPU

Headline workload (config.workload): BASELINE.json configs[1] — batch 64, 28 frames, 1536+1024 features,
hidden 512, vocab 10k, seq_len 30, greedy decode.  One "step" = one `SAModel.sample()` call on one batch
(Cross-Gating encoder + init state + 30 word steps).  `value` = captions/s with inputs resident in HBM;
`e2e` = the same call with pinned HOST inputs (H2D inside the timed region) and the token ids read back.
The train (config 3 / config 4) and beam (config 5) figures ride along in `train` / `beam` sub-objects.
One process per GPU; ranks work on independent batches (weak scaling, no collective on the decode path;
the train sub-object adds the one NCCL gradient allreduce per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = dict(R=1536, F=1024, H=512, E=468, A=1536, V=10000, C=14)
K_FRAMES, T_SEQ, BATCH = 28, 30, 64
METRIC = "captions/sec (train fwd+bwd; greedy decode) MSRVTT-shape batch @1/2/4/8 B200"   # BASELINE.json "metric", verbatim
L2_FLUSH_BYTES = 256 << 20
WORKLOAD = "config2: greedy decode, batch 64 per GPU, 28 frames, 1536+1024 feats, hidden 512, vocab 10k, seq_len 30"   # both arms
TRAIN_FLOP_PER_CAPTION = 3705.5e6      # algorithmic forward 1235.2 MFLOP (v2a once), train ~3x


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_opt(drop):
    return argparse.Namespace(vocab_size=DIMS["V"], category_size=DIMS["C"], input_encoding_size=DIMS["E"],
                              rnn_size=DIMS["H"], num_layers=1, drop_prob_lm=drop, seq_length=T_SEQ, seed=1024,
                              feat_size=DIMS["R"], feat_size2=DIMS["F"], att_size=DIMS["A"], fusion_activity="ReLU")


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference algorithm, all host threads)
# ----------------------------------------------------------------------------------------------
class CpuArm:
    """The reference's algorithm on the host cores: the reference's OWN files (baseline/_ref, staged by build() from
    /root/reference: kind "reference") when they travelled with the repo, else the oracle port (kind "port")."""

    def __init__(self, P):
        import torch
        from baseline import reference_arm as RA
        self.P = P
        self.kind = "port"
        self.RS = None
        if RA.available():
            try:
                self.RA = RA
                self.RS = RA.load()
                self.kind = "reference"
                self.model = RA.build_model(self.RS, DIMS, T_SEQ, 0.5, P)     # drop_prob_lm 0.5 as run_train.sh ships it
            except Exception as e:      # a torch too new for the shims: say so and time the port instead
                sys.stderr.write("[bench] staged reference not importable (%r): timing the oracle port\n" % (e,))
                self.RS, self.kind = None, "port"
        self.torch = torch

    def greedy(self, batch):
        if self.RS is not None:
            return self.RA.greedy(self.model, batch)
        from oracle import xgating_oracle as O
        with self.torch.no_grad():
            return O.sample_greedy(self.P, batch["rgb"], batch["opfl"], batch["feat_mask"], batch["pos"], T_SEQ)

    def train(self, batch):
        if self.RS is not None:
            return self.RA.train_step(self.model, self.RS, batch)
        from oracle import xgating_oracle as O
        return O.train_step_grads(self.P, batch, train=True)

    def beam(self, batch, rows):
        if self.RS is not None:
            return self.RA.beam(self.model, batch, 5, rows)
        from oracle import xgating_oracle as O
        sl = slice(0, rows)
        with self.torch.no_grad():
            V = O.encoder_fwd(self.P, batch["rgb"][sl], batch["opfl"][sl], batch["feat_mask"][sl])
            return O.sample_beam(self.P, V, batch["feat_mask"][sl], batch["pos"][sl], 5, T_SEQ)

    def seconds(self, fn, reps, warm=1):
        for _ in range(warm):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (all threads), same
    metric / config / warm-up policy as the GPU arm; each step = one batch of 64 captions."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import xgating_oracle as O          # input / weight generators (and the port when baseline/_ref is absent)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = O.synth_params(DIMS, 1024)
    P["logit.bias"][0] = -1e4
    batch = O.synth_inputs(DIMS, BATCH, K_FRAMES, T_SEQ, 0)
    arm = CpuArm(P)
    dt = arm.seconds(lambda: arm.greedy(batch), args.steps, warm=args.warmup)
    val = BATCH / dt
    tbatch = O.synth_inputs(DIMS, BATCH, K_FRAMES, T_SEQ, seed=100, full_length=False)
    tr = arm.seconds(lambda: arm.train(tbatch), max(1, min(args.steps, 3)), warm=1)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "captions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sample": "each step = one full batch of 64 captions on the host CPU, rank 0 only"},
            "cpu_baseline": {"value": val, "unit": "captions/s", "cores": cores, "kind": arm.kind,
                             "sample": "%d greedy batches of 64 captions (%s)" % (
                                 args.steps, "the reference's own caption_src files from baseline/_ref, torch CPU fp32"
                                 if arm.kind == "reference" else "oracle port of the reference, torch CPU fp32")},
            "train": {"value": BATCH / tr, "unit": "captions/s", "ms_per_step": tr * 1e3,
                      "workload": "config3: train fwd+bwd (XE loss), batch 64"},
            "e2e": {"value": val, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="xgating", choices=["xgating", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--skip-extra", action="store_true", help="skip the train / beam sub-objects")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the whole run (library
    # banners such as "NCCL version ..." are written to fd 1 from native code) and the line goes to the saved fd
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3)           # same warm-up policy in both arms
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the xgating path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line (NCCL's version banner goes to stderr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    import controllable_xgating_b200 as X
    import controllable_xgating_b200.SAModel as XS
    from controllable_xgating_b200 import _lib as XL
    from controllable_xgating_b200.parallel import DataParallelSAModel
    from oracle import xgating_oracle as O          # input / weight generators + the CPU baseline leg only
    XS.VERBOSE = False
    dev = torch.device("cuda", local)

    P = O.synth_params(DIMS, 1024)
    P["logit.bias"][0] = -1e4                       # EOS never chosen: every run executes all T steps
    batch = O.synth_inputs(DIMS, BATCH, K_FRAMES, T_SEQ, seed=rank)
    model = X.SAModel(make_opt(0.5))
    model.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    model.cuda().eval()
    model._engine.set_strict(True)                  # no silent per-step fallback on the measured path
    d = {k: v.to(dev) for k, v in batch.items() if isinstance(v, torch.Tensor)}
    pinned = {k: batch[k].pin_memory() for k in ("rgb", "opfl", "feat_mask", "pos", "seq", "seq_mask")}
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    gopt = {"sample_max": 1, "beam_size": 1}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """per-step CUDA events on the current stream; L2 flushed (256 MiB write) before every step, outside
        the timed interval; returns (total_ms over `steps`, max over ranks)."""
        for _ in range(warmup):
            flush.fill_(1)                          # warm-up steps look like the timed ones (the first step behind a flush
            fn()                                    # measured 0.2 ms slower than the following ones when they did not)
        barrier()
        tot = 0.0
        per_step = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
            per_step.append(e0.elapsed_time(e1))
        if os.environ.get("XG_BENCH_DEBUG") and rank == 0:
            print("[bench] per-step ms:", " ".join("%.3f" % x for x in per_step), file=sys.stderr)
        barrier()
        if world > 1:
            t = torch.tensor([tot], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot = float(t)
        return tot

    # ---- (1) greedy decode, inputs resident in HBM -------------------------------------------------
    steps_seen = []

    def greedy_resident():
        seq, _ = model.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], gopt)
        steps_seen.append(seq.shape[1])

    clocks = ClockSampler(local)
    clocks.start()
    ms = timed(greedy_resident, args.steps, args.warmup)
    value = world * BATCH * args.steps / (ms * 1e-3)

    # ---- (2) end to end through the public API: pinned host inputs, ids read back ---------------------
    # results are read back into pinned host buffers: both copies queued, ONE stream synchronisation
    host_seq = torch.empty((BATCH, T_SEQ), dtype=torch.int64).pin_memory()
    host_lp = torch.empty((BATCH, T_SEQ), dtype=torch.float32).pin_memory()

    def read_back(seq, lp):
        n = seq.shape[1]
        hs, hl = (host_seq, host_lp) if n == T_SEQ else (host_seq[:, :n], host_lp[:, :n])
        hs.copy_(seq, non_blocking=True); hl.copy_(lp, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return hs, hl

    def greedy_e2e():
        rgb = pinned["rgb"].to(dev, non_blocking=True); opfl = pinned["opfl"].to(dev, non_blocking=True)
        fm = pinned["feat_mask"].to(dev, non_blocking=True); pos = pinned["pos"].to(dev, non_blocking=True)
        seq, lp = model.sample(rgb, opfl, fm, pos, gopt)
        return read_back(seq, lp)

    ms_e2e_serial = timed(greedy_e2e, args.steps, args.warmup)

    # the same through the package's input pipeline (input_pipeline.DeviceStager): the host->device copy of batch
    # i+1 runs on a copy stream while batch i decodes.  One timed region over all K steps; every step's H2D, L2
    # flush, decode and id read-back are inside it.
    from controllable_xgating_b200.input_pipeline import DeviceStager
    stager = DeviceStager(dev)
    host_in = {k: pinned[k] for k in ("rgb", "opfl", "feat_mask", "pos")}

    # SAModel.sample_async queues a batch without the host synchronisation sample() needs (it has to learn how many
    # columns the reference would return): batch i+1 is queued before batch i is read back, so the GPU does not idle
    # while the host handles results.  Two pinned result buffers alternate.
    host_res = [(host_seq, host_lp), (torch.empty_like(host_seq).pin_memory(), torch.empty_like(host_lp).pin_memory())]
    from controllable_xgating_b200.SAModel import PendingSample

    def finish(pend, ev, slot):
        ev.synchronize()
        hs, hl = host_res[slot]
        n = PendingSample.steps_of(hs)
        return hs[:, :n], hl[:, :n]

    def greedy_pipelined(steps):
        out = []
        nxt = stager.put(**host_in)
        prev = None
        for i in range(steps):
            cur = nxt
            if i + 1 < steps:
                nxt = stager.put(**host_in)
            stager.wait(cur)
            flush.fill_(1)
            pend = model.sample_async(cur["rgb"], cur["opfl"], cur["feat_mask"], cur["pos"], gopt)
            pend.to_host(*host_res[i & 1])                  # D2H of ids + log-probs queued behind the decode
            ev = torch.cuda.Event(); ev.record()
            if prev is not None:
                out.append(finish(*prev))                   # the previous batch: wait for ITS copies only
            prev = (pend, ev, i & 1)
        if prev is not None:
            out.append(finish(*prev))
        return out

    greedy_pipelined(args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    greedy_pipelined(args.steps)
    e1.record(); e1.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    barrier()
    if world > 1:
        tt = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt)
    clk = clocks.stop()
    e2e_value = world * BATCH * args.steps / (ms_e2e * 1e-3)
    h2d = sum(pinned[k].numel() * pinned[k].element_size() for k in ("rgb", "opfl", "feat_mask", "pos"))
    d2h = BATCH * T_SEQ * (8 + 4)

    # ---- (3) per-kernel timing pass (CUDA events around every launch, same workload) ------------------
    eng = model._engine
    lib = XL.load()
    XL.check(lib.xg_profile_enable(eng.handle, 1), "xg_profile_enable", eng.handle)
    import ctypes
    buf = ctypes.create_string_buffer(1 << 20)

    def prof_round():
        flush.fill_(1)
        greedy_resident()
        XL.check(lib.xg_profile_report(eng.handle, buf, len(buf)), "xg_profile_report", eng.handle)   # drains the records
        return {p["name"]: p for p in json.loads(buf.value.decode())}

    prof_round()                                    # untimed: creates the event pool, first launch with events on
    prof_steps = 7
    rounds = [prof_round() for _ in range(prof_steps)]
    XL.check(lib.xg_profile_enable(eng.handle, 0), "xg_profile_enable", eng.handle)
    # one sample per kernel per round (ms summed over that round's launches / launches); the per-kernel figure is the
    # MEDIAN round, so that a single host hiccup between the two events of one launch does not move it (all samples
    # are printed next to it)
    names = sorted({n for r in rounds for n in r})
    prof = []
    for n in names:
        per = sorted(r[n]["ms"] / r[n]["launches"] for r in rounds if n in r)
        launches = sum(r[n]["launches"] for r in rounds if n in r)
        med = per[len(per) // 2]
        prof.append({"name": n, "launches": launches, "ms": med * launches, "samples_us": [round(x * 1e3, 1) for x in per],
                     "mean_us": sum(per) / len(per) * 1e3})
    launches_per_step = sum(p["launches"] for p in prof) / prof_steps
    prof.sort(key=lambda p: -p["ms"])
    tot_prof_ms = sum(p["ms"] for p in prof)
    top = prof[0]
    peak, peak_src = peaks()

    def algorithmic_bytes(name):
        """DESIGN.md section 4 / SURVEY.md 8(d): bytes a launch must move once, whatever the implementation."""
        H, E, A, V, R, F = DIMS["H"], DIMS["E"], DIMS["A"], DIMS["V"], DIMS["R"], DIMS["F"]
        if name == "decode_persistent":
            # per word step: every decoder weight once (52.5 MB fp32 incl. the logit matrix) + V and Uv of every caption
            w = 4.0 * (A * 2 * H + H * E + 4 * H * E + 5 * 4 * H * H + V * H)
            act = 4.0 * BATCH * K_FRAMES * (H + A)
            return T_SEQ * (w + act)
        if name == "encode_persistent":
            # per frame step (27 of them): both recurrent matrices once (16.8 MB with the input halves hoisted: 8.4 MB)
            return (K_FRAMES - 1) * 4.0 * 2 * 4 * H * H
        # a GEMM launch must read both operands once and write C once
        if name.startswith("gemm"):
            m, n, k = (int(x) for x in name.split("_")[2].split("x"))
            return 4.0 * (m * k + n * k + m * n)
        if name == "att_fwd":      # AH + Uv + V read, alpha + context written, per caption row
            return 4.0 * BATCH * (A + K_FRAMES * A + K_FRAMES * H + K_FRAMES + H)
        return None

    def measured_traffic(name):
        """dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel (per launch),
        recorded under profiles/ by the round's profiling run; None if no capture is committed."""
        pth = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(pth):
            try:
                return json.load(open(pth)).get(name, {}).get("dram_bytes_per_launch")
            except Exception:
                return None
        return None

    ab = algorithmic_bytes(top["name"])
    avg_ms = top["ms"] / top["launches"]
    roofline = {"kernel": top["name"], "bound": "hbm", "achieved": (ab / (avg_ms * 1e-3) / 1e9) if ab else None,
                "peak": peak, "unit": "GB/s", "frac": (ab / (avg_ms * 1e-3) / 1e9 / peak) if ab else None,
                "traffic": measured_traffic(top["name"]), "peak_source": peak_src, "algorithmic_bytes_per_launch": ab,
                "avg_launch_us": avg_ms * 1e3, "launch_us_stat": "median of %d profiled steps (1 launch each)" % prof_steps,
                "launch_us_samples": top["samples_us"], "launch_us_mean": top["mean_us"],
                "share_of_step": top["ms"] / tot_prof_ms,
                "top5": [{"name": p["name"], "launches_per_step": p["launches"] / prof_steps,
                          "us_per_launch": p["ms"] / p["launches"] * 1e3, "share": p["ms"] / tot_prof_ms} for p in prof[:5]]}

    line = {"metric": METRIC, "value": value, "unit": "captions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "word_steps_executed": int(min(steps_seen)) if steps_seen else None,
                       "l2": "flushed (256 MiB write) before every timed step", "weights": "random init (reference init distributions)",
                       "parallelism": "dp%d (independent batches, no collective on the decode path)" % world},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "captions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "how": "DeviceStager double buffering (H2D of step i+1 on a copy stream under the decode of step i) + "
                           "SAModel.sample_async (step i+1 is queued before the ids of step i are read back); "
                           "one timed region over all steps with every step's H2D, L2 flush, decode and D2H inside",
                    "unpipelined_value": world * BATCH * args.steps / (ms_e2e_serial * 1e-3),
                    "unpipelined_ms_per_step": ms_e2e_serial / args.steps},
            "gpu_launches": int(round(launches_per_step * args.steps)),
            "roofline": roofline}

    # ---- (4) train fwd+bwd (config 3 at N=1; config 4 = global batch 512 at N>1) + beam-5 (config 5) ------
    if not args.skip_extra:
        tb = BATCH if world == 1 else 512 // world
        tbatch = O.synth_inputs(DIMS, tb, K_FRAMES, T_SEQ, seed=100 + rank, full_length=False)
        td = {k: v.to(dev) for k, v in tbatch.items() if isinstance(v, torch.Tensor)}
        tp = {k: tbatch[k].pin_memory() for k in ("rgb", "opfl", "feat_mask", "pos", "seq", "seq_mask")}
        tmodel = X.SAModel(make_opt(0.5))
        tmodel.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
        tmodel.cuda().train()
        tmodel._engine.set_strict(True)
        crit = X.LanguageModelCriterion()
        dp = DataParallelSAModel(tmodel, overlap=os.environ.get("XG_DP_OVERLAP", "1") != "0") if world > 1 else None
        fwd = dp if dp is not None else tmodel

        def train_step(src):
            for p_ in tmodel.parameters():
                p_.grad = None
            logp, _ = fwd(src["rgb"], src["opfl"], src["feat_mask"], src["pos"], src["seq"], src["seq_mask"])
            loss = crit(logp, src["seq"], src["seq_mask"])
            loss.backward()
            return loss

        ms_t = timed(lambda: train_step(td), max(3, args.steps // 2), args.warmup)
        nst = max(3, args.steps // 2)

        def train_e2e():
            src = {k: v.to(dev, non_blocking=True) for k, v in tp.items()}
            return float(train_step(src))
        ms_te = timed(train_e2e, nst, 1)
        line["train"] = {"workload": ("config3: train fwd+bwd (XE loss, dropout 0.5), batch 64" if world == 1 else
                                      "config4: train fwd+bwd, global batch 512 (%d per GPU), one NCCL gradient allreduce per step" % tb),
                         "value": world * tb * nst / (ms_t * 1e-3), "unit": "captions/s", "ms_per_step": ms_t / nst,
                         "e2e": {"value": world * tb * nst / (ms_te * 1e-3), "unit": "captions/s",
                                 "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in tp.values()),
                                 "d2h_bytes_per_step": 4},
                         "allreduce_bytes_per_step": (dp.hook.bytes // max(dp.hook.calls, 1)) if dp else 0}
        # the train step is dense-contraction work: its roofline is the tensor pipe.  fp32-grade results come from three
        # tf32 MMAs per product (3xTF32), so the reachable ceiling of THIS arithmetic is peak / 6 (tf32 = bf16 / 2, x 1/3).
        tpk, tpk_src = tensor_peak()
        tfl = TRAIN_FLOP_PER_CAPTION * tb * world
        ach = tfl / (ms_t / nst * 1e-3) / 1e12
        line["train"]["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tpk * world, "unit": "TFLOP/s", "frac": ach / (tpk * world),
                                     "peak_source": tpk_src, "algorithmic_flops_per_step": tfl,
                                     "cap_3xtf32": "fp32-grade accuracy costs 3 tf32 MMAs per product: ceiling = peak/6 = %.0f TFLOP/s"
                                                   % (tpk * world / 6.0),
                                     "frac_of_3xtf32_cap": ach / (tpk * world / 6.0)}
        if world == 1:
            # the reference's whole iteration ("time/batch", starttrain.py:125-137): forward, backward, elementwise
            # gradient clamp +-0.1 and Adam, all derived parameter tables rebuilt every step.  Timed LAST on this model:
            # it moves the weights
            from controllable_xgating_b200.optim import FusedAdam
            fopt = FusedAdam(tmodel.parameters(), lr=4e-4, grad_clip=0.1)

            def train_iter():
                train_step(td)
                fopt.step()
            ms_it = timed(train_iter, nst, 2)
            line["train"]["with_optimizer"] = {"workload": "config3 + clamp(+-0.1) + Adam step (xg_adam_step), parameter-derived tables rebuilt",
                                               "value": tb * nst / (ms_it * 1e-3), "unit": "captions/s", "ms_per_step": ms_it / nst}
            bopt = {"beam_size": 5}
            ms_b = timed(lambda: model.sample(d["rgb"], d["opfl"], d["feat_mask"], d["pos"], bopt), 3, 1)
            line["beam"] = {"workload": "config5: sample_beam beam_size 5, batch 64", "value": BATCH * 3 / (ms_b * 1e-3),
                            "unit": "captions/s", "ms_per_step": ms_b / 3}

    # ---- (5) CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample) --------------------
    if rank == 0 and world == 1 and not args.skip_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        arm = CpuArm(P)
        reps = 10
        sec = arm.seconds(lambda: arm.greedy(batch), reps)
        src = "the reference's own caption_src files (baseline/_ref)" if arm.kind == "reference" else "the oracle port of the reference"
        line["cpu_baseline"] = {"value": BATCH / sec, "unit": "captions/s", "cores": cores, "kind": arm.kind,
                                "sample": "%d greedy batches of 64 captions (config 2) with %s, torch CPU fp32, %d threads"
                                          % (reps, src, cores)}
        if not args.skip_extra:
            sect = arm.seconds(lambda: arm.train(tbatch), 3)
            line["cpu_baseline"]["train_value"] = BATCH / sect
            rows = 4
            secb = arm.seconds(lambda: arm.beam(batch, rows), 1, warm=0)
            line["cpu_baseline"]["beam_value"] = rows / secb
            line["cpu_baseline"]["beam_sample"] = "beam-5 on the first %d videos of the batch (config 5), one pass" % rows
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
