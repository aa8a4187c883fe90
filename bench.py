#!/usr/bin/env python3
"""bench.py -- HE-AAC synthesis path throughput on B200 (the BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA back end
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU decoder (oracle/_ref)

A "step" is one pass of the hot path over one batch: `frames` consecutive frames of `streams`
independent streams per GPU (weak scaling: per-GPU work is fixed; streams are sharded across ranks with no
collective on the data path -- NCCL only gathers the timings).  Workload at N=1 = BASELINE config 2
(HE-AAC v1 stereo, 24 kHz core + SBR -> 48 kHz, 1024 streams).

  value     decoded audio-seconds per second, inputs already resident in HBM, timed with CUDA events on
            the launching stream, max over ranks
  e2e       same metric through the C-ABI with HOST buffers (pinned), H2D/D2H inside the timed region
  roofline  dominant kernel: algorithmic bytes per launch / its average launch time vs measured HBM peak
  cpu_baseline  the reference decoder (full decode, parse included) on all host cores, bounded sample
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

import numpy as np  # noqa: E402

# algorithmic bytes per unit (s16 output, tables cache-resident)
ALG_BYTES = {"lc_chframe": 10240, "v1_chframe": 38912, "v2_streamframe": 66392}
# of which each kernel's share when the stages run as separate launches (state it owns + its I/O); DESIGN.md "kernels"
KERNEL_ALG_BYTES = {
    "k_core":  4096 + 2 * 2048,                      # coeffs + saved R/W  (+ PCM 2048 when it is the last stage)
    # k_chain, whole HE-AAC v1 channel-frame: coeffs + params + state R+W + PCM; when it stops at X (PS /
    # downsampled SBR) it has neither the synthesis history nor the PCM
    "k_chain": 38912,
    "k_chain_x": 25600,
    "k_ps":    1024 + 13144,
    "k_synth": 2 * 4608 + 4096,                      # v history R+W + PCM, per output channel
}

# dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed `ncu --set full` captures
# (profiles/r01_ncu_summary.md, prof_r1d): keyed by (kernel, workload, streams per GPU).
NCU_DRAM_BYTES_PER_LAUNCH = {("k_sbr", "heaac_v1_stereo", 1024): 33.671680e6 + 6.095872e6,
                             ("k_core", "heaac_v1_stereo", 1024): 15.283712e6 + 0.025856e6}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:                                             # NVML: sub-millisecond queries, many samples per timed region
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            while not self._halt.is_set():
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for b, n in bits.items():
                    if r & b:
                        self.reasons.add(n)
                self._halt.wait(0.005)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(o[0])); self.max_mhz = float(o[1])
                for n, v in zip(names, o[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def audio_seconds(workload, S, T):
    from ffmpeg_heaac_b200 import synthetic
    w = synthetic.WORKLOADS[workload]
    rate = {3: 48000, 4: 44100, 6: 24000}[w["sr_index"]]
    return S * T * 1024.0 / rate


def units_per_frame(workload, S):
    """(unit name, algorithmic bytes per unit, units per frame tick)"""
    from ffmpeg_heaac_b200 import synthetic
    w = synthetic.WORKLOADS[workload]
    ncc = {"mono": 1, "stereo": 2, "5.1": 6}[w["layout"]]
    if w["ps"]:
        return "stream-frame", ALG_BYTES["v2_streamframe"], S
    if w["sbr"]:
        return "channel-frame", ALG_BYTES["v1_chframe"], S * ncc
    return "channel-frame", ALG_BYTES["lc_chframe"], S * ncc


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(workload, seconds_budget, n_threads=None, n_streams=None, n_frames=48):
    """The reference's own decoder (oracle/_ref: unmodified aacdec/aacsbr/aacps, full decode incl. parsing),
    one decoder context per host thread, all threads busy.  Returns dict(value, cores, sample, ...)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyref
    import heaac_streams as hs
    from ffmpeg_heaac_b200 import synthetic
    w = synthetic.WORKLOADS[workload]
    cores = n_threads or os.cpu_count() or 1
    n_streams = n_streams or min(cores, 16)
    pyref.lib().ref_set_log_level(-8)
    decs, pkts_all, rate_out = [], [], None
    for i in range(cores):
        seed = 1000 + (i % n_streams)
        if w["sbr"]:
            stereo = w["layout"] == "stereo"
            pkts, ed, info = hs.make_stream(seed, n_frames, stereo=stereo, ps=w["ps"])
            d = pyref.Decoder(ed, 24000, 2 if stereo else 1, "stereo" if stereo else "mono", float_mode=False)
            rate_out = 48000
        else:
            ch = {"mono": 1, "stereo": 2, "5.1": 6}[w["layout"]]
            rate = {3: 48000, 4: 44100}[w["sr_index"]]
            pcm = pyref.lcg_noise(1024 * (n_frames + 1), ch, seed=seed)
            enc = pyref.Encoder(rate, ch, 64000 * ch)
            pkts = enc.encode(pcm)[:n_frames]
            ed = enc.extradata
            enc.close()
            d = pyref.Decoder(ed, rate, ch, w["layout"], float_mode=False)
            rate_out = rate
        decs.append(d)
        pkts_all.append(pkts)
    # calibrate on one thread, then run all threads for ~seconds_budget
    t0 = time.perf_counter()
    decs[0].decode_many(pkts_all[0], 1)
    per_pass = max(time.perf_counter() - t0, 1e-4)
    repeat = max(1, int(seconds_budget / per_pass))
    results = [0] * cores

    def work(i):
        results[i] = decs[i].decode_many(pkts_all[i], repeat)

    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    if any(r < 0 for r in results):
        raise RuntimeError("reference decode failed in the CPU baseline")
    audio = sum(results) / float(rate_out)
    for d in decs:
        d.close()
    return {"value": audio / dt, "unit": "audio-s/s", "cores": cores, "kind": "reference",
            "sample": "%d decoder contexts x %d passes over %d-frame %s streams (%.1f s wall, full decode incl. parse)"
                      % (cores, repeat, n_frames, workload, dt), "seconds": dt}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="heaac_v1_stereo")
    ap.add_argument("--streams", type=int, default=None, help="streams per GPU (default: the BASELINE batch)")
    ap.add_argument("--frames", type=int, default=32, help="frames per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    default_streams = {"heaac_v1_stereo": 1024, "heaac_v2_ps": 4096, "lc_51_48k_tns": 2048, "lc_stereo_44k": 4096}
    S = args.streams or default_streams[args.workload]
    T = args.frames
    cfg_desc = {"workload": args.workload, "streams_per_gpu": S, "frames_per_step": T,
                "layout": None, "l2": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(args.workload, seconds_budget=max(2.0, args.cpu_seconds))
        line = {"impl": "reference", "metric": "decoded audio-seconds per second (batched streams)", "value": r["value"],
                "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg_desc,
                "cpu_baseline": {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import ffmpeg_heaac_b200 as hb
    from ffmpeg_heaac_b200 import synthetic

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    torch.cuda.set_device(dev)

    cfg, a = synthetic.make_workload(args.workload, S, T, seed=1234 + rank, K=32)
    cfg.device = dev
    cfg.want_float_out = 0
    cfg_desc["layout"] = synthetic.WORKLOADS[args.workload]["layout"]
    ctx = hb.Context(cfg)

    # ---- device-resident inputs (the `value` leg) ----
    bufs = {}
    for k in ("coeffs", "ics", "tns", "sbr", "ps"):
        if a[k] is not None:
            bufs[k] = ctx.alloc(a[k].nbytes).upload(a[k])
    pcm_bytes = T * S * ctx.out_len * ctx.nout * 2
    d_pcm = ctx.alloc(pcm_bytes)
    in_bytes = sum(v.nbytes for v in a.values() if v is not None)

    def step():
        ctx.synth_device(T, bufs["coeffs"], bufs["ics"], d_pcm, tns=bufs.get("tns"), sbr=bufs.get("sbr"), ps=bufs.get("ps"))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(dev)
    sampler.start()
    l0 = ctx.launch_count
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()           # CUDA events on the launching stream
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.launch_count - l0
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
    else:
        ms_max = ms
    audio = audio_seconds(args.workload, S, T) * args.steps * world
    value = audio / (ms_max * 1e-3)

    # ---- per-kernel timing pass (same work, events around every launch) ----
    ctx.profile_enable(4 * T * 4 + 16)
    prof = {}
    nprof = min(args.steps, 4)
    for _ in range(nprof):
        step()
        for k, (tms, n) in ctx.profile_read().items():
            p = prof.setdefault(k, [0.0, 0])
            p[0] += tms; p[1] += n
    ctx.profile_enable(0)
    peak, peak_kind = load_peaks()
    unit_name, unit_bytes, units = units_per_frame(args.workload, S)
    kern = {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] // nprof} for k, v in prof.items() if v[1]}
    tot = sum(v["ms_per_launch"] for v in kern.values())
    dom = max(kern, key=lambda k: kern[k]["ms_per_launch"])
    per_launch_units = S if dom == "k_ps" else (S * ctx.nout if dom == "k_synth" else S * ctx.ncc)
    if dom == "k_chain":
        per_launch_units *= kern[dom]["launches_per_step"] == 1 and T or 1      # one launch walks through all T frames (v1)
    dom_bytes = KERNEL_ALG_BYTES["k_chain_x" if (dom == "k_chain" and cfg.ps) else dom] + (2048 if (dom == "k_core" and not cfg.sbr) else 0)
    achieved = per_launch_units * dom_bytes / (kern[dom]["ms_per_launch"] * 1e-3) / 1e9
    path_achieved = units * unit_bytes * T * args.steps / (ms_max * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((dom, args.workload, S)), "peak_kind": peak_kind, "kernel_share_of_step": kern[dom]["ms_per_launch"] / tot,
                "kernels": kern,
                # per-kernel times come from the serialised pass (an event pair around every launch); in the timed
                # region the kernels of a tick are chained with programmatic dependent launches and their last waves
                # overlap, so the step is shorter than launches x the sum of these
                "kernel_timing": "serialised pass", "serialised_ms_per_step": tot * T,
                "path": {"achieved": path_achieved, "frac": path_achieved / peak, "alg_bytes_per_unit": unit_bytes, "unit_name": unit_name}}

    # ---- e2e leg: host (pinned) buffers through the C-ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        pinned = {}
        for k, v in a.items():
            if v is None:
                pinned[k] = None
                continue
            t = torch.empty(v.nbytes, dtype=torch.uint8).pin_memory()
            arr = t.numpy().view(v.dtype).reshape(v.shape)
            arr[...] = v
            pinned[k] = (t, arr)
        out_t = torch.empty(pcm_bytes, dtype=torch.uint8).pin_memory()
        out = out_t.numpy().view(np.int16).reshape(T, S, ctx.out_len, ctx.nout)
        g = lambda k: pinned[k][1] if pinned[k] is not None else None

        def e2e_step():
            ctx.synth(g("coeffs"), g("ics"), g("tns"), g("sbr"), g("ps"), pcm_out=out)
            return int(out[0, 0, 0, 0])          # the step's result is read on the host

        e2e_step()
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": audio_seconds(args.workload, S, T) * n_e2e * world / dt, "unit": "audio-s/s",
               "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(pcm_bytes), "steps": n_e2e}

    # ---- packets leg: bitstreams in host memory -> PCM in host memory through heaac_dec_decode (own host
    #      parser on all host threads + the GPU back end); like-for-like with the reference arm's full decode ----
    e2e_packets = e2e_boundary = e2e_host_parser = None
    fixture = os.path.join(ROOT, "tests", "golden", "bench_%s_packets.npz" % args.workload)
    if not args.no_e2e and os.path.exists(fixture):
        z = np.load(fixture)
        sizes, blob = z["pkt_sizes"], z["pkt_blob"].tobytes()
        K, N = sizes.shape
        ofs = np.concatenate([[0], np.cumsum(sizes.ravel())])
        pk = [[blob[ofs[k * N + i]:ofs[k * N + i + 1]] for i in range(N)] for k in range(K)]
        nb = max(1, N // T)
        batches = [hb.PacketBatch([[pk[(s + rank) % K][(b * T + t) % N] for s in range(S)] for t in range(T)]) for b in range(nb)]
        out_t2 = torch.empty(pcm_bytes, dtype=torch.uint8).pin_memory()
        out2 = out_t2.numpy().view(np.int16).reshape(T, S, ctx.out_len, ctx.nout)
        out_t3 = torch.empty(pcm_bytes, dtype=torch.uint8).pin_memory()
        out3 = out_t3.numpy().view(np.int16).reshape(T, S, ctx.out_len, ctx.nout)

        def packets_run(gpu_parse):
            dec = hb.Decoder(S, z["extradata"].tobytes(), int(z["core_rate"]), int(z["channels"]), max_frames=T, device=dev,
                             gpu_parse=gpu_parse)
            outs = [out2, out3]
            state = {"i": 0, "prev": None}

            def pkt_step():
                """submit step i, then collect step i-1: two calls in flight, every step's PCM read on the host"""
                i = state["i"]
                tk = dec.submit(batches[i % nb], outs[i % 2])
                r = 0
                if state["prev"] is not None:
                    dec.wait(state["prev"])
                    r = int(outs[(i - 1) % 2][0, 0, 0, 0])
                state["prev"], state["i"] = tk, i + 1
                return r

            def drain():
                if state["prev"] is not None:
                    dec.wait(state["prev"])
                    state["prev"] = None
                return int(outs[(state["i"] - 1) % 2][0, 0, 0, 0])

            dec.decode(batches[0], pcm_out=out2, out_shape=(ctx.out_len, ctx.nout))     # configures the decoder
            state["i"] = 1
            pkt_step()
            drain()
            barrier()
            n_pk = max(2, min(args.steps, 16))
            l0p = dec.info()["gpu_launches"]
            t0 = time.perf_counter()
            for _ in range(n_pk):
                pkt_step()
            drain()
            dtp = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dtp], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtp = float(t.item())
            inf = dec.info()
            dec.close()
            return {"value": audio_seconds(args.workload, S, T) * n_pk * world / dtp, "unit": "audio-s/s",
                    "h2d_bytes_per_step": int(batches[0].nbytes) if gpu_parse else int(in_bytes),
                    "d2h_bytes_per_step": int(pcm_bytes), "steps": n_pk,
                    "parser": "gpu (k_parse: one warp per stream, the frames of a chunk across its lanes)" if gpu_parse else "host (%d threads)" % (os.cpu_count() or 1),
                    "parse_errors": inf["parse_errors"], "gpu_launches": int(inf["gpu_launches"] - l0p),
                    "k_parse_ms_per_launch": inf["k_parse_ms"] if gpu_parse else None, "calls_in_flight": 2,
                    "input": "%d looped fixture streams x %d frames replicated over %d streams" % (K, N, S)}

        e2e_packets = packets_run(True)
        e2e_packets["host_parser"] = packets_run(False)
        # The packets -> PCM call is the whole decode, like the reference arm's: it is the headline `e2e`; the
        # coefficient-level call at the spectral_to_sample boundary is kept beside it.
        e2e_boundary, e2e = e2e, dict(e2e_packets)
        e2e_host_parser = e2e.pop("host_parser")
        e2e["api"] = "heaac_dec_submit / heaac_dec_wait: packets in host memory -> s16 PCM in host memory"
        if e2e_boundary is not None:
            e2e_boundary["api"] = "heaac_b200_synth: parsed coefficients + side info in host memory -> s16 PCM in host memory"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(args.workload, seconds_budget=args.cpu_seconds)
            cpu = {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "audio-s/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}

    if rank == 0:
        line = {"metric": "decoded audio-seconds per second (batched streams)", "value": value, "unit": "audio-s/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg_desc, "clocks": clocks, "e2e": e2e, "e2e_boundary": e2e_boundary,
                "e2e_host_parser": e2e_host_parser, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "wall_s_timed_region": wall}
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
