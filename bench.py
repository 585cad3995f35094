#!/usr/bin/env python
"""bench.py -- W-HMR regressor-loop hot path on N B200s (BASELINE.json configs[1]).

One step = one pass of the body-model hot path of WHMR.forward's loop over a batch of B=256 bodies
per GPU (whmr_b200.loop.RegressorLoop.step): init SMPL, 3 x {MAF sampling, SMPL, read-outs, weak +
predicted-focal projection}, global SMPL -- 5 SMPL forwards, 3 samplings, 7 projections per body.
metric = SMPL bodies/s (batch rows through the whole loop), whole job over all N GPUs.

  value     : inputs resident in HBM, the step replayed as one CUDA graph, CUDA-event timed.
  e2e       : through the host-buffer path -- every step copies ALL step inputs (feature maps 4.2 GB
              + parameters) from pinned host memory, runs the step, and copies the results the
              reference's caller reads (demo/tester.py:164-165) back to pinned host memory; H2D, compute
              and D2H run on three streams over two buffer sets (step k+1's copies under step k's kernels).
  roofline  : dominant kernel CLASS of the step (the three sampling launches are one class), timed inside
              the timed region with external CUDA events recorded as graph nodes; algorithmic bytes/flops
              per DESIGN.md, DRAM bytes actually moved from the committed ncu capture.
  channels_last / other_configs : the same step with channels_last feature maps (NHWC kernel); BASELINE
              configs[2] (65,536-body SMPL + H36M sweep point, strong scaling over the ranks) and configs[4]
              (35,515-frame evaluation pass + NCCL gather) -- so the multi-GPU record has a collective in it.
  cpu_baseline : the oracle (CPU restatement of the reference's path, torch CPU kernels, all host
              threads) on a bounded sample of the same workload (rank 0, N=1 only).
`--impl reference` times that CPU path alone, same metric/config.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "smpl_bodies_per_sec"
UNIT = "bodies/s"
V, J, NBETA, KPOSE = 6890, 24, 10, 207
VP = 6912


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.stop_flag, self.region = index, period, [], False, "idle"
        self.ok = False
        self.pcie = {}
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((self.region, sm, int(r)))
                if self.region.startswith("pcie:"):     # end-to-end legs: what really crosses the link (20 ms counter windows)
                    rx = nv.nvmlDeviceGetPcieThroughput(self.h, nv.NVML_PCIE_UTIL_RX_BYTES)   # KB/s, host -> device
                    tx = nv.nvmlDeviceGetPcieThroughput(self.h, nv.NVML_PCIE_UTIL_TX_BYTES)
                    self.pcie.setdefault(self.region[5:], []).append((rx, tx))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def pcie_summary(self, tag, ms_per_step):
        """median NVML PCIe counters over the leg `tag` -> measured GB/s and bytes per step (None without samples)"""
        v = self.pcie.get(tag) or []
        if len(v) < 3:
            return None
        rx = sorted(x[0] for x in v)[len(v) // 2] * 1024.0
        tx = sorted(x[1] for x in v)[len(v) // 2] * 1024.0
        return {"rx_GBs": rx / 1e9, "tx_GBs": tx / 1e9, "rx_bytes_per_step": rx * ms_per_step * 1e-3,
                "tx_bytes_per_step": tx * ms_per_step * 1e-3, "samples": len(v),
                "source": "nvmlDeviceGetPcieThroughput (median of 20 ms windows during the timed steps, rank 0's GPU)"}

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvml unavailable: %s" % getattr(self, "err", "")}
        load = [s for s in self.samples if s[0] != "idle"]
        timed = [s for s in load if s[0] == "timed"] or load
        sm = sorted(s[1] for s in timed)
        reasons = set()
        for s in load:
            for bit, name in self.REASONS.items():
                if s[2] & bit:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(reasons),
                "samples_under_load": len(load), "samples_timed": len([s for s in load if s[0] == "timed"])}


def bind_to_gpu_numa_node(dev_index):
    """Pin this process (and so its first-touch pinned allocations) to the CPUs of the NUMA node the GPU hangs off:
    eight ranks pulling GBs per step through pinned buffers on the wrong socket halve their H2D rate.  Best effort."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(dev_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return {"node": None, "note": "no NUMA affinity reported for %s" % bus}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return {"node": node, "note": "node CPUs not in this process's affinity mask"}
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed), "pci": bus}
    except Exception as e:  # noqa: BLE001
        return {"node": None, "note": "%s: %s" % (type(e).__name__, str(e)[:80])}


# ------------------------------------------------------------------------------------------------
# algorithmic bytes / flops per launch (DESIGN.md section "Kernels")
# ------------------------------------------------------------------------------------------------
def algorithmic(kind, B, ctx):
    C = 256
    if kind == "chain":
        return {"bytes": B * (4 * (NBETA + 216) + 4 * (J * 12 + J * 3) + ctx["pf_bytes"]), "bound": "hbm"}
    if kind == "pose_blend":
        return {"flops": B * 2.0 * KPOSE * 3 * V, "bound": "tensor"}
    if kind == "blend_skin":   # smpl_fused_tc: pose+shape blend GEMM, transform blend and skinning in one kernel
        return {"flops": B * 2.0 * KPOSE * 3 * V, "bytes": B * (4 * 3 * V + 4 * J * 12 + 2 * 2 * 224), "bound": "tensor"}
    if kind == "skin":
        return {"bytes": B * (4 * 3 * VP + 4 * 3 * V + 4 * J * 12 + 4 * NBETA), "bound": "hbm"}
    if kind == "readout":
        return {"bytes": B * 12 * (ctx["readout_nnz"] + ctx["readout_rows"]), "bound": "hbm"}
    if kind == "skin_readout":   # skinning kernel (with the one-hot read-outs in its epilogue) + regressor rows
        return {"bytes": B * (4 * 3 * VP + 4 * 3 * V + 4 * J * 12 + 4 * NBETA + 12 * (ctx["readout_nnz"] + ctx["readout_rows"])),
                "bound": "hbm"}
    if kind.startswith("sample_l"):
        lvl = int(kind[-1])
        H, W = ctx["levels"][lvl]
        N = 63 if lvl == 0 else 67
        return {"bytes": B * (4 * C * (min(4 * N, H * W) + N) + 8 * N), "bound": "hbm"}
    if kind == "project_weak_full":
        return {"bytes": B * (4 * 7 * 49 + 12 + 36 + 16), "bound": "hbm"}
    if kind == "project_weak":
        return {"bytes": B * (4 * 5 * 49 + 12), "bound": "hbm"}
    if kind == "project_markers":
        return {"bytes": B * (4 * 5 * 67 + 12), "bound": "hbm"}
    if kind == "project_full":
        return {"bytes": B * (4 * 5 * 49 + 36 + 16), "bound": "hbm"}
    return {"bytes": 0, "bound": "hbm"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import whmr_b200.synthetic as syn
    from whmr_b200 import _lib, ops
    from whmr_b200.loop import RegressorLoop, make_loop_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    gemm_mode = args.gemm_mode

    sampler = ClockSampler(local_rank)
    sampler.start()

    model = syn.make_smpl_model(seed=0, weights="random")
    loop = RegressorLoop(model, dev, backbone=args.backbone, gemm_mode=gemm_mode)
    feats, params, bbox = make_loop_inputs(B, dev, backbone=args.backbone, seed=1, rank=rank)
    if args.channels_last:
        feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]

    # ---- parity gate on this rank's own data before anything is timed (16 bodies vs the oracle) ----
    parity = None
    if rank == 0 and not args.skip_parity:
        from oracle.loop_oracle import LoopOracle, to_cpu_inputs
        n = 16
        sub = ([f[:n].contiguous() for f in feats], [{k: v[:n].contiguous() for k, v in q.items()} for q in params],
               {k: v[:n].contiguous() for k, v in bbox.items()})
        got = loop.step(*sub)
        torch.cuda.synchronize()
        ref = LoopOracle(model, args.backbone).step(*to_cpu_inputs(*sub))
        e_v = float((got["verts"].cpu() - ref["verts"]).abs().max())
        e_g = float((got["global_verts"].cpu() - ref["global_verts"]).abs().max())
        e_k = float(((got["kp_2d_w"].cpu() - ref["kp_2d_w"]).abs() * (sub[2]["orig_shape"].cpu()[:, [1, 0]] / 2).unsqueeze(1)).max())
        e_f = max(float((a.cpu() - b).abs().max() / b.abs().max()) for a, b in zip(got["point_feats"], ref["point_feats"]))
        parity = {"verts_m": max(e_v, e_g), "kp2d_px": e_k, "sampled_rel": e_f, "bodies": n}
        if not (parity["verts_m"] <= 1e-5 and e_f <= 1e-4 and e_k <= 1e-3):   # north-star tolerances
            raise SystemExit("bench.py: parity gate failed: %s" % parity)

    # ---- device-resident timing: the step as one CUDA graph ----------------------------------------
    n0 = _lib.launch_count()
    loop.step(feats, params, bbox)
    launches_per_step = _lib.launch_count() - n0
    graph, outs = loop.capture(feats, params, bbox)
    for _ in range(W):
        graph.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.region = "timed"
    ev0.record()
    for _ in range(K):
        graph.replay()
    ev1.record()
    torch.cuda.synchronize()
    sampler.region = "load"
    if world > 1:
        dist.barrier()
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total) / K
    value = world * B / (ms_step * 1e-3)

    # ---- per-kernel timing inside the (instrumented) graph, same K steps ---------------------------
    marks = []

    def probe(name):
        e = torch.cuda.Event(enable_timing=True, external=True)
        if name == "pre_smpl":
            a = torch.cuda.Event(enable_timing=True, external=True)
            b = torch.cuda.Event(enable_timing=True, external=True)
            c = torch.cuda.Event(enable_timing=True, external=True)
            h.set_probe_events(a, b, c)
            e.record()
            marks.append(("start", e))
            marks.append(("chain", a))
            marks.append(("pose_blend", b))
            marks.append(("skin", c))
        else:
            e.record()
            marks.append(("readout" if name == "skin_readout" else name, e))

    h, _ = loop.smpl._state(dev)
    kern = {}
    if rank == 0:
        loop.head.probe = probe
        marks.clear()
        igraph, _ = loop.capture(feats, params, bbox, warmup=0)
        loop.head.probe = None
        h.set_probe_events(None, None, None)
        mk = list(marks)
        for _ in range(3):
            igraph.replay()
        torch.cuda.synchronize()
        acc = {}
        reps = min(K, 200)
        for _ in range(reps):
            igraph.replay()
            torch.cuda.synchronize()
            for (n_prev, e_prev), (name, e) in zip(mk[:-1], mk[1:]):
                if name == "start":
                    continue
                t = e_prev.elapsed_time(e)
                a = acc.setdefault(name, [0.0, 0])
                a[0] += t
                a[1] += 1
        if h.is_fused() and "pose_blend" in acc and "skin" in acc:
            # one kernel: the pose_blend probe sits right after the chain kernel, the skin probe after the fused kernel
            gap, sk = acc.pop("pose_blend"), acc.pop("skin")
            acc["blend_skin"] = [gap[0] + sk[0], sk[1]]
        ro = loop.head._readout(dev, loop.with_h36m)
        ctx = {"pf_bytes": 2 * 208 * (2 if loop.smpl.gemm_mode == ops.GEMM_TC_BF16X3 else 4),
               "readout_nnz": int(ro.csr.nnz), "readout_rows": int(ro.R), "levels": loop.levels}
        for name, (tot, cnt) in acc.items():
            per_launch_ms = tot / cnt
            launches = cnt // reps
            a = algorithmic(name, B, ctx)
            k = {"launches_per_step": launches, "ms_per_launch": per_launch_ms, "ms_per_step": per_launch_ms * launches,
                 "bound": a["bound"]}
            if a["bound"] == "tensor":
                mode_peak = peaks["bf16_tflops_sustained"] / 3.0 if loop.smpl.gemm_mode == ops.GEMM_TC_BF16X3 \
                    else peaks["bf16_tflops_sustained"] / 2.0 / 3.0
                k.update(achieved=a["flops"] / (per_launch_ms * 1e-3) / 1e12, peak=mode_peak, unit="TFLOP/s")
                if "bytes" in a:   # the same launch against the HBM roofline of its output stream
                    k["hbm_achieved_GBs"] = a["bytes"] / (per_launch_ms * 1e-3) / 1e9
                    k["hbm_frac"] = k["hbm_achieved_GBs"] / peaks["hbm_gbs"]
            else:
                k.update(achieved=a["bytes"] / (per_launch_ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s")
            k["frac"] = k["achieved"] / k["peak"] if k["peak"] else None
            tr = ncu_traffic(name, B)   # bytes the kernel actually moved (committed ncu capture), against the live time
            # (only when the capture has the same launch count per pass as this probed schedule: the deferred schedule
            #  batches the five read-out finishes into one launch)
            if tr is not None and a["bound"] == "hbm" and tr.get("launches") in (None, launches):
                k["dram_bytes_per_launch_ncu"] = tr["dram_bytes_per_launch"]
                k["dram_GBs_moved"] = tr["dram_bytes_per_launch"] / (per_launch_ms * 1e-3) / 1e9
                k["dram_frac_moved"] = k["dram_GBs_moved"] / peaks["hbm_gbs"]
            kern[name] = k
        lv = [kern[n] for n in ("sample_l0", "sample_l1", "sample_l2") if n in kern]
        if len(lv) == 3:   # the three sampling launches of a step as ONE kernel class
            alg = sum(algorithmic(n, B, ctx)["bytes"] for n in ("sample_l0", "sample_l1", "sample_l2"))
            ms = sum(k["ms_per_step"] for k in lv)
            agg = {"launches_per_step": 3, "ms_per_launch": ms / 3, "ms_per_step": ms, "bound": "hbm",
                   "achieved": alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "algorithmic_bytes_per_step": alg}
            agg["frac"] = agg["achieved"] / agg["peak"]
            if all("dram_bytes_per_launch_ncu" in k for k in lv):
                moved = sum(k["dram_bytes_per_launch_ncu"] for k in lv)
                agg.update(dram_bytes_per_step_ncu=moved, dram_GBs_moved=moved / (ms * 1e-3) / 1e9,
                           dram_frac_moved=moved / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], wasted_traffic_factor=moved / alg)
            kern["sampling"] = agg

    # ---- the same step with channels_last feature maps (NHWC sampling kernel; no copy: a backbone run in channels_last
    #      hands over exactly this memory) -----------------------------------------------------------------------------
    cl = None
    if not args.skip_channels_last and not args.channels_last:
        feats_cl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
        g_cl, outs_cl = loop.capture(feats_cl, params, bbox)
        for _ in range(W):
            g_cl.replay()
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        e_pf = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs_cl["point_feats"], outs["point_feats"]))
        same_verts = bool(torch.equal(outs_cl["verts"], outs["verts"]))
        if world > 1:
            dist.barrier()
        ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kc = min(K, 300)
        ca.record()
        for _ in range(kc):
            g_cl.replay()
        cb.record()
        torch.cuda.synchronize()
        t_cl = torch.tensor([ca.elapsed_time(cb)], device=dev)
        if world > 1:
            dist.all_reduce(t_cl, op=dist.ReduceOp.MAX)
        ms_cl = float(t_cl) / kc
        cl = {"ms_per_step": ms_cl, "value": world * B / (ms_cl * 1e-3), "unit": UNIT, "steps": kc,
              "parity_vs_nchw": {"sampled_rel": e_pf, "verts_identical": same_verts},
              "note": "feature maps in torch.channels_last memory format: the NHWC sampling kernel reads 128 contiguous bytes "
                      "per tap instead of one DRAM atom per (tap row, channel)"}
        if not (e_pf <= 1e-4 and same_verts):
            raise SystemExit("bench.py: channels_last leg differs from the NCHW step: %s" % cl["parity_vs_nchw"])
        del g_cl, outs_cl, feats_cl
        torch.cuda.empty_cache()

    # ---- the step INCLUDING the extractors' reduce_dim MLP: sampling + projection + Conv1d MLP as ONE
    #      launch per level (maf_fused_kernel, 3xTF32 on tcgen05) against the two-step path (sampling kernel, [B,256,N]
    #      round trip through HBM, PyTorch Conv1d MLP) --------------------------------------------------------------------
    rd = None
    if not args.skip_reduce_dim:
        rd = reduce_dim_leg(args, model, dev, feats, params, bbox, rank, world, min(K, 300), W, gemm_mode)

    # ---- training step: Regressor.forward body-model head forward + loss + backward at the reference's
    #      TRAIN.BATCH_SIZE, custom ops with their CUDA backward kernels vs eager autograd of the dense path ----------------
    train = None
    if rank == 0 and not args.skip_train:
        train = train_step_leg(model, loop, dev)

    # ---- end to end through host buffers ------------------------------------------------------------
    e2e = None
    e2e_copy = None
    e2e_gather = None
    e2e_gather_cl = None
    e2e_resident = None
    if not args.skip_e2e:
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)  # noqa: E731
        h_feats = [pin(f) for f in feats]
        h_params = [{k: pin(v) for k, v in q.items()} for q in params]
        h_bbox = {k: pin(v) for k, v in bbox.items()}
        out_keys = ["verts", "global_verts", "pred_cam_t", "focal_length", "kp_2d_w", "global_kp_3d", "theta"]
        h2d_small = sum(v.numel() * 4 for q in h_params for v in q.values()) + sum(v.numel() * 4 for v in h_bbox.values())
        h2d_feat = sum(f.numel() * 4 for f in h_feats)
        s_h2d, s_d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        s_cmp = torch.cuda.current_stream(dev)

        depth = max(2, int(args.e2e_depth))

        cmp_streams = [s_cmp]
        cmp_loops = [loop]

        def make_sets(own_feats, host_levels=(), n_cs=1):
            """`depth` independent buffer sets (inputs, captured graph, outputs, pinned result buffers).  Levels in
            `host_levels` stay in pinned host memory: the sampling kernel gathers their taps in place over PCIe.
            n_cs > 1: consecutive steps replay on n_cs compute streams, each with its own loop object (SMPL handle and
            workspace), so the PCIe gathers of two steps overlap."""
            while len(cmp_streams) < n_cs:
                cmp_streams.append(torch.cuda.Stream(device=dev))
                cmp_loops.append(RegressorLoop(model, dev, backbone=args.backbone, gemm_mode=gemm_mode))
            sets = []
            for i in range(depth if n_cs == 1 else max(depth, 2 * n_cs) // n_cs * n_cs):
                f_i = [torch.empty_like(f) for f in feats] if own_feats else list(feats)
                p_i = [{k: v.clone() for k, v in q.items()} for q in params]
                b_i = {k: v.clone() for k, v in bbox.items()}
                if own_feats:
                    for d_, s_ in zip(f_i, feats):
                        d_.copy_(s_)
                for lv in host_levels:
                    f_i[lv] = h_feats[lv]
                g_i, o_i = cmp_loops[i % n_cs].capture(f_i, p_i, b_i)
                h_o = {k: torch.empty(o_i[k].shape, dtype=o_i[k].dtype, pin_memory=True) for k in out_keys}
                sets.append({"feats": f_i, "params": p_i, "bbox": b_i, "graph": g_i, "outs": o_i, "h_out": h_o,
                             "cmp": cmp_streams[i % n_cs],
                             "ev_in": torch.cuda.Event(), "ev_cmp": torch.cuda.Event(), "ev_out": torch.cuda.Event()})
            return sets

        def run_pipeline(sets, with_feats, steps):
            """step k: H2D of its inputs (stream 1) -> graph replay (stream 2) -> D2H of its results (stream 3); buffer set
            k % depth, so the copies of step k+1 run under the kernels of step k and the read-back of step k-1.  The host
            waits for the results of step k-(depth-1) before it enqueues step k+1 (a caller that consumes every result,
            depth-1 steps behind the one it is submitting)."""
            for st in sets:
                st["ev_cmp"].record(st["cmp"])
                st["ev_out"].record(s_d2h)
            D = len(sets)
            for k in range(steps):
                st = sets[k % D]
                with torch.cuda.stream(s_h2d):
                    s_h2d.wait_event(st["ev_cmp"])          # the previous replay on this set has consumed its inputs
                    if with_feats:
                        for d_, s_ in zip(st["feats"], h_feats):
                            if d_ is not s_:                # a host-resident level is read in place by the sampling kernel
                                d_.copy_(s_, non_blocking=True)
                    for dq, sq in zip(st["params"], h_params):
                        for kk in dq:
                            dq[kk].copy_(sq[kk], non_blocking=True)
                    for kk in st["bbox"]:
                        st["bbox"][kk].copy_(h_bbox[kk], non_blocking=True)
                    st["ev_in"].record(s_h2d)
                s_c = st["cmp"]
                s_c.wait_event(st["ev_in"])
                s_c.wait_event(st["ev_out"])                # the previous results of this set have left the device
                with torch.cuda.stream(s_c):
                    st["graph"].replay()
                st["ev_cmp"].record(s_c)
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(st["ev_cmp"])
                    for kk in out_keys:
                        st["h_out"][kk].copy_(st["outs"][kk], non_blocking=True)
                    st["ev_out"].record(s_d2h)
                if k >= D - 1:
                    sets[(k - (D - 1)) % D]["ev_out"].synchronize()   # the caller reads step k-(D-1)'s results now
            for j in range(max(0, steps - (D - 1)), steps):
                sets[j % D]["ev_out"].synchronize()

        def time_e2e(sets, with_feats, steps, tag=None, _again=False):
            run_pipeline(sets, with_feats, 3)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            if tag:
                if _again:
                    sampler.pcie.pop(tag, None)
                sampler.region = "pcie:" + tag
            t0 = time.perf_counter()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(s_cmp)
            run_pipeline(sets, with_feats, steps)
            torch.cuda.synchronize()
            b_.record(s_cmp)
            torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3
            sampler.region = "load"
            t = torch.tensor([max(a_.elapsed_time(b_), wall_ms)], device=dev)   # the slower of device and host clocks
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if tag and float(t) < 600.0 and not _again:    # long enough for the 20 ms PCIe counter windows (same on all ranks)
                return time_e2e(sets, with_feats, int(steps * 700.0 / max(float(t), 1.0)) + 1, tag, True)
            return float(t) / steps

        ke = max(4, min(K, args.e2e_steps))
        sets = make_sets(True)
        d2h = sum(v.numel() * 4 for v in sets[0]["h_out"].values())
        ms_e2e = time_e2e(sets, True, ke, "copy_all")
        del sets
        torch.cuda.empty_cache()
        # the same pass with the sparsely sampled levels left in pinned host memory: the sampling kernel gathers their taps
        # in place (unified addressing), so only the sectors the taps touch cross PCIe instead of the whole map
        host_levels = [int(x) for x in args.e2e_host_levels.split(",") if x.strip() != ""]
        if host_levels:
            n_cs = max(1, int(args.e2e_compute_streams))
            sets = make_sets(True, host_levels, n_cs)
            ms_g = time_e2e(sets, True, ke, "host_gather")
            sets_n = list(range(len(sets)))
            pf_g = [t.clone() for t in sets[-1]["outs"]["point_feats"]]
            v_g = sets[-1]["outs"]["verts"].clone()
            del sets
            torch.cuda.empty_cache()
            same = all(torch.equal(a_, b_) for a_, b_ in zip(pf_g, outs["point_feats"])) and torch.equal(v_g, outs["verts"])
            if not same:
                raise SystemExit("bench.py: the pass with host-resident feature levels differs from the device-resident step")
            copied = h2d_small + sum(h_feats[l].numel() * 4 for l in range(len(h_feats)) if l not in host_levels)
            n_pts = [outs["point_feats"][l].shape[-1] for l in range(len(h_feats))]
            gathered = sum(B * feats[l].shape[1] * n_pts[l] * 2 * 32 for l in host_levels)
            e2e_gather = {"value": world * B / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g, "steps": ke,
                          "h2d_bytes_per_step": copied + gathered, "h2d_bytes_copied": copied,
                          "h2d_bytes_gathered_in_place": gathered, "host_input_bytes_per_step": h2d_small + h2d_feat,
                          "d2h_bytes_per_step": d2h, "host_resident_levels": host_levels,
                          "pipeline": "H2D + %d compute + D2H streams x %d buffer sets" % (n_cs, len(sets_n)), "results_identical_to_device_resident_step": True,
                          "note": "ALL step inputs in pinned host memory; levels %s (%s) are NOT copied: the sampling kernel reads "
                                  "their taps in place through unified addressing (gathered bytes = 2 tap rows x one 32-byte "
                                  "sector per point and channel); the other levels and all parameters are cudaMemcpyAsync'd"
                                  % (host_levels, ", ".join("%dx%d" % tuple(feats[l].shape[2:]) for l in host_levels))}
        # ... and with channels_last host maps (a backbone run in channels_last), every level gathered in place: a tap is then
        # 1 KB of contiguous channels, so all three levels together move ~0.2 GB over PCIe
        e2e_gather_cl = None
        if host_levels and not args.skip_channels_last and not args.channels_last:
            h_feats.clear()        # the NCHW host copies are not needed again: their pinned blocks are reused below
            for f in feats:
                hb = torch.empty((f.shape[0], f.shape[2], f.shape[3], f.shape[1]), dtype=f.dtype, pin_memory=True)
                hb.copy_(f.permute(0, 2, 3, 1))
                h_feats.append(hb.permute(0, 3, 1, 2))        # [B,C,H,W] view in channels_last strides, still pinned
            sets = make_sets(False, list(range(len(feats))))
            ms_gc = time_e2e(sets, True, max(ke, min(K, 50)), "host_gather_cl")
            e_cl = max(float((a_ - b_).abs().max() / b_.abs().max()) for a_, b_ in zip(sets[0]["outs"]["point_feats"], outs["point_feats"]))
            del sets
            torch.cuda.empty_cache()
            if not e_cl <= 1e-4:
                raise SystemExit("bench.py: channels_last host-gather pass differs from the NCHW step: %g" % e_cl)
            gathered = sum(B * f.shape[1] * n_pts[l] * 4 * 4 for l, f in enumerate(feats))
            e2e_gather_cl = {"value": world * B / (ms_gc * 1e-3), "unit": UNIT, "ms_per_step": ms_gc,
                             "h2d_bytes_per_step": h2d_small + gathered, "h2d_bytes_copied": h2d_small,
                             "h2d_bytes_gathered_in_place": gathered, "host_input_bytes_per_step": h2d_small + h2d_feat,
                             "d2h_bytes_per_step": d2h, "sampled_rel_vs_nchw_step": e_cl,
                             "note": "all three levels in pinned host memory in channels_last format, none copied: 4 taps x "
                                     "C contiguous floats per point (not the reference's NCHW layout: reported beside the "
                                     "headline, not as it)"}
        sets = make_sets(False)
        ms_e2e_res = time_e2e(sets, False, max(ke, min(K, 200)), "feat_resident")
        del sets
        torch.cuda.empty_cache()
        e2e_copy = {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_small + h2d_feat,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": ke, "pipeline": "3 streams x %d buffer sets" % depth,
               "h2d_GBs_per_rank": (h2d_small + h2d_feat) / (ms_e2e * 1e-3) / 1e9, "numa": numa,
               "note": "ALL step inputs from pinned host memory, incl. the 3 feature-map levels (in the reference "
                       "these are produced on the device by the backbone and never cross PCIe)"}
        # headline: the faster way of feeding the SAME host-resident inputs through the public API
        e2e = e2e_gather if (e2e_gather and e2e_gather["value"] > e2e_copy["value"]) else e2e_copy
        e2e_resident = {"value": world * B / (ms_e2e_res * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_small,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e_res, "pipeline": "3 streams x %d buffer sets" % depth,
                        "d2h_GBs_per_rank": d2h / (ms_e2e_res * 1e-3) / 1e9,
                        "note": "same, feature maps device-resident as in the reference (backbone output)"}
        for d_, tag_ in ((e2e_copy, "copy_all"), (e2e_gather, "host_gather"), (e2e_gather_cl, "host_gather_cl"),
                         (e2e_resident, "feat_resident")):
            if d_ is not None:
                d_["pcie_measured"] = sampler.pcie_summary(tag_, d_["ms_per_step"])
    sampler.region = "idle"

    # ---- kernel quality at scale: one SMPL forward at 16k bodies (BASELINE configs[2]) -------------
    scale = None
    if rank == 0 and not args.skip_sweep:
        scale = smpl_sweep(loop, dev, peaks, [4096, 16384] if not args.quick else [4096])

    modes = None
    if rank == 0 and not args.skip_sweep and gemm_mode is None:
        modes = pose_blend_modes_leg(loop, model, dev, feats, params, bbox, B, peaks, reps=min(K, 300))

    # ---- BASELINE configs[2] and configs[4] on all ranks: a strong-scaling point and a pass that ends in a collective ----
    other = None
    if not args.skip_other:
        other = other_configs(loop, model, dev, rank, world, peaks)

    # ---- the path's share of the whole regressor loop, old vs new, rank 0, N == 1 ----------------
    whole = None
    if rank == 0 and world == 1 and not args.skip_whole_loop:
        whole = whole_loop_leg(model, args.backbone, dev, B)

    # ---- CPU baseline (oracle on host cores), rank 0, N == 1 ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_loop_baseline(args.backbone, args.cpu_sample or 64, reps=3)

    # ---- the same dense PyTorch path, eager, on THIS GPU ("the real bar"), rank 0, N == 1 ---------------
    eager = None
    if rank == 0 and world == 1 and not args.skip_eager:
        try:
            eager = torch_gpu_eager_baseline(model, args.backbone, feats, params, bbox, B)
        except Exception as e:  # noqa: BLE001  (a baseline leg must never take the bench line down)
            eager = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}

    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    clocks = sampler.summary()

    if rank == 0:
        classes = {n: k for n, k in kern.items() if not n.startswith("sample_l")} if "sampling" in kern else kern
        dom = max(classes.items(), key=lambda kv: kv[1]["ms_per_step"])[0] if classes else None
        roof = None
        if dom:
            k = kern[dom]
            total_ms = sum(x["ms_per_step"] for x in classes.values())
            if dom == "sampling":
                traffic = ({"dram_bytes_per_launch": k["dram_bytes_per_step_ncu"] / 3, "source": "profiles/traffic.json (ncu --set full), mean of the 3 launches"}
                           if "dram_bytes_per_step_ncu" in k else None)
            else:
                traffic = ncu_traffic(dom, B)
            roof = {"kernel": dom, "bound": k["bound"], "achieved": k["achieved"], "peak": k["peak"], "unit": k["unit"],
                    "frac": k["frac"], "traffic": traffic, "peak_source": peaks["source"] + (
                        " (sustained bf16 / 3 MMAs per product)" if k["bound"] == "tensor" else " hbm copy"),
                    "ms_per_launch": k["ms_per_launch"], "launches_per_step": k["launches_per_step"],
                    "share_of_step": k["ms_per_step"] / total_ms}
            if "dram_frac_moved" in k:
                roof.update(frac_of_bytes_moved=k["dram_frac_moved"], wasted_traffic_factor=k.get("wasted_traffic_factor"))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "whmr_regressor_loop_B256 (BASELINE configs[1]): init SMPL + 3 x (MAF sampling + SMPL + "
                                   "read-outs + weak/full projection) + global SMPL; %s feature levels %s x 256 ch"
                                   % (args.backbone, list(loop.levels)),
                       "batch_per_gpu": B, "global_batch": B * world, "smpl_forwards_per_step": 5,
                       "samplings_per_step": 3, "pose_blend_arithmetic": {0: "fp32_simt", 1: "tcgen05 bf16x3", 2: "tcgen05 3xtf32"}[loop.smpl.gemm_mode],
                       "parallelism": "dp%d (bodies sharded by rank, no data-path collective)" % world,
                       "l2": "inputs larger than L2: feature maps 4.2 GB/step/GPU + ~0.2 GB of outputs vs 126 MB L2",
                       "launch": "one CUDA graph replay per step; schedule: finishing passes of the 5 read-outs + the 4 joint projections after the loop in ONE launch, per-kernel probes taken on the immediate 22-launch schedule",
                       "feature_layout": "channels_last (NHWC memory)" if args.channels_last else "NCHW contiguous (reference layout)",
                       "rotation_glue": "unbiased_gram_schmidt (eval mode) + rotation_matrix_to_angle_axis + theta inside the chain kernel, every SMPL call"},
            "clocks": clocks, "e2e": e2e, "e2e_copy_all": e2e_copy, "e2e_host_gather": e2e_gather, "e2e_host_gather_channels_last": e2e_gather_cl, "e2e_feat_resident": e2e_resident, "channels_last": cl, "with_reduce_dim": rd, "train_step": train, "whole_loop": whole, "other_configs": other,
            "gpu_launches": int(launches_per_step) * K, "gpu_launches_per_step": int(launches_per_step),
            "roofline": roof, "kernels": kern, "cpu_baseline": cpu, "torch_gpu_eager": eager, "smpl_at_scale": scale, "pose_blend_modes": modes, "parity": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def train_step_leg(model, loop, dev, B=64, reps=30):
    """core/trainer.py:380-636 back-propagates through the SMPL vertices, the 49 / H36M joints and both projections of
    Regressor.forward (cfg.TRAIN.BATCH_SIZE = 64, configs/pymaf_config.yaml:28).  One head forward + a loss over those
    tensors + backward: the torch.library ops of this repo (whmr_smpl_backward, whmr_readout_backward,
    whmr_project_*_backward) against eager PyTorch autograd of the same dense path (the oracle restatement on cuda:0),
    gradients w.r.t. rotmat / betas / cam / Tz compared between the two."""
    import torch
    import whmr_b200.synthetic as syn
    from oracle.loop_oracle import LoopOracle
    b = syn.make_bodies(B, seed=5)
    T = lambda a, g=False: torch.from_numpy(a).to(dev).requires_grad_(g)  # noqa: E731
    rm, be, cam, tz = T(b["rotmat"], True), T(b["betas"], True), T(b["cam"], True), T(b["Tz"], True)
    bbox = {"bbox_height": T(b["bbox_height"]), "center": T(b["center"]), "orig_shape": T(b["orig_shape"]), "Tz": tz}
    head = loop.head
    keys = ("verts", "kp_3d", "kp_2d", "kp_2d_w", "smpl_kp_3d")
    stage_before = head.train_stage
    head.train_stage = 2          # both projections carry the joint gradient somewhere (models/whmr.py:143-173)

    def loss_of(o):
        return o["verts"].pow(2).sum() + o["kp_3d"].pow(2).sum() + o["kp_2d"].pow(2).sum() + o["kp_2d_w"].pow(2).sum() + \
            o["smpl_kp_3d"].pow(2).sum()

    def ours():
        o = head(rm, be, cam, bbox["bbox_height"], bbox["center"], bbox["orig_shape"], tz, J_regressor=True, is_train=True)
        loss_of(o).backward()

    orc = LoopOracle(model, loop.backbone, device=str(dev))

    def eager():
        o = orc.regressor_outputs({"rotmat": rm, "betas": be, "cam": cam}, bbox, is_train=True, train_stage=2)
        loss_of(o).backward()

    def grads(fn):
        for t in (rm, be, cam, tz):
            t.grad = None
        fn()
        return [t.grad.clone() if t.grad is not None else torch.zeros_like(t) for t in (rm, be, cam, tz)]

    def timed(fn):
        for _ in range(5):
            grads(fn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    g_ours, g_ref = grads(ours), grads(eager)
    rel = [float((a - r_).abs().max() / r_.abs().max().clamp_min(1e-30)) for a, r_ in zip(g_ours, g_ref)]
    ms_ours, ms_eager = timed(ours), timed(eager)

    def graph_ms(fn):
        """the same forward + loss + backward captured once as a CUDA graph and replayed: device time without the host's
        per-op dispatch (both arms are host-bound at B = 64 when launched eagerly)"""
        try:
            for t in (rm, be, cam, tz):
                t.grad = None
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            for t in (rm, be, cam, tz):
                t.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        except Exception as e:  # noqa: BLE001
            torch.cuda.synchronize()
            return "capture failed: %s" % (str(e).splitlines()[0][:160],)

    ms_ours_graph, ms_eager_graph = graph_ms(ours), graph_ms(eager)
    head.train_stage = stage_before
    for t in (rm, be, cam, tz):
        t.grad = None
    return {"batch": B, "ms_forward_loss_backward": ms_ours, "ms_eager_autograd_same_gpu": ms_eager,
            "speedup_vs_eager": ms_eager / ms_ours,
            "ms_graph_replay": ms_ours_graph, "ms_eager_autograd_graph_replay": ms_eager_graph,
            "grad_rel_diff_vs_eager": dict(zip(("rotmat", "betas", "cam", "Tz"), rel)),
            "loss": "sum of squares of verts, kp_3d (H36M), kp_2d, kp_2d_w, smpl_kp_3d; train stage 2 detach routing",
            "note": "ms_forward_loss_backward / ms_eager_autograd_same_gpu: host-launched (eager) on both sides; "
                    "ms_graph_replay: the same forward + loss + backward of this repo's ops captured as one CUDA graph "
                    "(the dense reference path builds tensors on the host inside the step and cannot be captured); backward "
                    "kernels are CUDA-core (FFMA) kernels"}


def reduce_dim_leg(args, model, dev, feats, params, bbox, rank, world, K, W, gemm_mode):
    """The loop step with the three MAF_Extractor modules in it (random-init Conv1d MLPs 256->128->64->32 with skip concats,
    models/maf_extractor.py:75-101): fused (one maf_fused_kernel launch per level, no [B,256,N] tensor) vs two-step
    (sampling kernel -> HBM -> PyTorch MLP), NCHW and channels_last maps; parity of the 'ref_features' of 16 bodies
    against the oracle (grid_sample + reduce_dim on the CPU, fp32) at <= 1e-4 relative."""
    import torch
    import torch.distributed as dist
    from whmr_b200 import _lib
    from whmr_b200.loop import RegressorLoop
    from whmr_b200.maf_extractor import MAF_Extractor
    torch.manual_seed(1234)
    exts = [MAF_Extractor(mesh_downsampling=None) for _ in range(3)]
    convs = [[(c.weight.detach().clone(), c.bias.detach().clone()) for c in e.filters] for e in exts]
    loop = RegressorLoop(model, dev, backbone=args.backbone, gemm_mode=gemm_mode, extractors=exts)
    for e in loop.extractors:
        e.return_point_feat = False
    B = feats[0].shape[0]
    res = {"mlp": "Conv1d(k=1) 256->128, [128;256]->64, [64;256]->32, leaky_relu x2 + relu, random init",
           "arithmetic_fused": "3xTF32 (tcgen05 kind::tf32, hi/lo split of both operands)", "unit": UNIT}
    if rank == 0 and not args.skip_parity:
        from oracle.loop_oracle import LoopOracle, to_cpu_inputs
        n = 16
        sub = ([f[:n].contiguous() for f in feats], [{k: v[:n].contiguous() for k, v in q.items()} for q in params],
               {k: v[:n].contiguous() for k, v in bbox.items()})
        got = loop.step(*sub)
        torch.cuda.synchronize()
        ref = LoopOracle(model, args.backbone, convs=convs).step(*to_cpu_inputs(*sub))
        err = max(float((a.cpu() - b).abs().max() / b.abs().max()) for a, b in zip(got["ref_features"], ref["ref_features"]))
        res["parity_ref_features_rel"] = err
        if not err <= 1e-4:
            raise SystemExit("bench.py: fused sampling + reduce_dim differs from the oracle: %.3e relative" % err)

    def timed(fs, fused):
        for e in loop.extractors:
            e.fused = fused
        n0 = _lib.launch_count()
        loop.step(fs, params, bbox)
        ours = _lib.launch_count() - n0
        g, _ = loop.capture(fs, params, bbox)
        for _ in range(W):
            g.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t) / K
        del g
        return {"ms_per_step": ms, "value": world * B / (ms * 1e-3), "our_launches_per_step": int(ours)}

    feats_cl = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    res["fused_nchw"] = timed(feats, True)
    res["two_step_nchw"] = timed(feats, False)
    res["fused_channels_last"] = timed(feats_cl, True)
    res["two_step_channels_last"] = timed(feats_cl, False)
    res["two_step_note"] = ("sampling kernel + the module's PyTorch reduce_dim (cuDNN/cuBLAS convs, cat, leaky_relu: ~10 "
                            "library launches per level, not counted in our_launches_per_step)")
    del feats_cl
    torch.cuda.empty_cache()
    return res


def ncu_traffic(kernel, B):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at batch B, from the committed
    `ncu --set full` capture (profiles/traffic.json, written by tools/ncu_traffic.py); None if not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    try:
        t = json.load(open(p)).get("B%d" % B, {}).get(kernel)
        return None if t is None else {"dram_bytes_per_launch": t["dram_bytes"], "source": t["source"],
                                       "launches": t.get("launches")}
    except Exception:  # noqa: BLE001
        return None


def smpl_sweep(loop, dev, peaks, sizes):
    """SMPL forward alone at large batch: per-kernel CUDA-event times via the stage entry points."""
    import torch
    import whmr_b200.synthetic as syn
    from whmr_b200 import ops
    h, _ = loop.smpl._state(dev)
    res = {}
    for Bs in sizes:
        b = syn.make_bodies(Bs, seed=5)
        betas = torch.from_numpy(b["betas"]).to(dev)
        rot = torch.from_numpy(b["rotmat"]).to(dev)
        for _ in range(3):
            h.forward(betas, rot, True)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 10
        e[0].record()
        for _ in range(reps):
            h.forward(betas, rot, True)
        e[1].record()
        torch.cuda.synchronize()
        ms = e[0].elapsed_time(e[1]) / reps
        bytes_alg = Bs * (4 * (NBETA + 216) + 4 * (3 * V + 3 * J))
        gb = bytes_alg / (ms * 1e-3) / 1e9
        mode = loop.smpl.gemm_mode
        tpeak = peaks["bf16_tflops_sustained"] / (3.0 if mode == ops.GEMM_TC_BF16X3 else 6.0)
        tf = Bs * 2.0 * KPOSE * 3 * V / (ms * 1e-3) / 1e12
        res[str(Bs)] = {"ms": ms, "bodies_per_s": Bs / (ms * 1e-3), "hbm_GBs_algorithmic": gb,
                        "hbm_frac": gb / peaks["hbm_gbs"], "pose_blend_TFLOPs_algorithmic_if_alone": tf,
                        "tensor_frac_lower_bound": tf / tpeak}
    return res


def pose_blend_modes_leg(loop, model, dev, feats, params, bbox, B, peaks, reps=300):
    """The two tensor-core arithmetics of the fused SMPL kernel side by side (the north star names TF32 / 3xTF32; bf16x3 is
    the production default): loop step time at this batch, SMPL alone at 16,384 bodies against that arithmetic's own tensor
    ceiling (sustained bf16 / 3 for bf16x3, sustained tf32 = bf16 / 2, / 3 for 3xTF32), and the measured vertex error
    against the fp64 oracle.  Rank 0."""
    import torch
    import whmr_b200.synthetic as syn
    from oracle.smpl_oracle import SMPLOracle
    from whmr_b200 import ops
    keep = loop.smpl.gemm_mode
    orc = SMPLOracle(model, torch.float64)
    bb = syn.make_bodies(32, seed=77)
    ref = orc(bb["betas"], bb["rotmat"][:, 1:], bb["rotmat"][:, :1], pose2rot=False)["vertices"]
    big = syn.make_bodies(16384, seed=5)
    betas, rot = torch.from_numpy(big["betas"]).to(dev), torch.from_numpy(big["rotmat"]).to(dev)
    res = {}
    for name, mode, div in (("bf16x3", ops.GEMM_TC_BF16X3, 3.0), ("3xtf32", ops.GEMM_TC_3XTF32, 6.0)):
        loop.smpl.set_gemm_mode(mode)
        h, _ = loop.smpl._state(dev)
        v = h.forward(torch.from_numpy(bb["betas"]).to(dev), torch.from_numpy(bb["rotmat"]).to(dev), True)[0]
        err = float((v.double().cpu() - ref).abs().max())
        g, _o = loop.capture(feats, params, bbox)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        z.record()
        torch.cuda.synchronize()
        ms_step = a.elapsed_time(z) / reps
        del g, _o
        for _ in range(3):
            h.forward(betas, rot, True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            h.forward(betas, rot, True)
        z.record()
        torch.cuda.synchronize()
        ms_big = a.elapsed_time(z) / 10
        bps = 16384 / (ms_big * 1e-3)
        tf = bps * 2.0 * KPOSE * 3 * V / 1e12
        res[name] = {"fused_kernel": bool(h.is_fused()), "ms_per_step": ms_step, "step_bodies_per_s": B / (ms_step * 1e-3),
                     "smpl_16384_bodies_per_s": bps, "pose_blend_TFLOPs_algorithmic": tf,
                     "tensor_ceiling_TFLOPs": peaks["bf16_tflops_sustained"] / div,
                     "tensor_frac": tf / (peaks["bf16_tflops_sustained"] / div), "verts_max_err_vs_fp64_m": err}
    loop.smpl.set_gemm_mode(keep)
    loop.smpl._state(dev)
    return res


def other_configs(loop, model, dev, rank, world, peaks):
    """BASELINE configs[2] at its largest point (65,536 bodies in total: SMPL + H36M joint regression, sharded over the
    ranks = strong scaling) and configs[4] (35,515 frames: GT SMPL + predicted SMPL + H36M 17->14 + MPJPE / PA-MPJPE / PVE,
    sharded, then ONE all_gather_into_tensor of the [n,3] errors).  Timed on the device, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import whmr_b200.synthetic as syn
    from whmr_b200.dist import all_gather_rows, shard_bounds
    from whmr_b200.evaluate import EvalPass

    def timed(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ev = EvalPass(loop.smpl, model["J_regressor_h36m"])
    res = {}
    # ---- configs[0]: ONE SMPL forward at B = 64, axis-angle and rotation-matrix input (latency-bound: launch + one pass over
    #      the posedirs tiles); rank 0, next to the CPU oracle on the host cores
    if rank == 0:
        from whmr_b200 import ops
        b64 = syn.make_bodies(64, seed=11)
        h, _ = loop.smpl._state(dev)
        be, aa, rm = (torch.from_numpy(b64[k]).to(dev) for k in ("betas", "pose_aa", "rotmat"))
        row = {"workload": "configs[0]: one SMPL forward (vertices + 24 joints), batch 64"}
        for tag, pose, is_rot in (("axis_angle", aa, False), ("rotmat", rm, True)):
            for _ in range(3):
                h.forward(be, pose, is_rot)
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(200):
                h.forward(be, pose, is_rot)
            b_.record()
            torch.cuda.synchronize()
            us = a_.elapsed_time(b_) / 200 * 1e3
            row[tag] = {"us_per_forward_eager": us, "bodies_per_s": 64 / (us * 1e-6)}
        if world == 1:
            from oracle.smpl_oracle import SMPLOracle
            use_all_host_threads()
            orc = SMPLOracle(model)
            import time as _t
            orc(b64["betas"], b64["pose_aa"][:, 3:], b64["pose_aa"][:, :3])
            ts = []
            for _ in range(5):
                t0 = _t.perf_counter()
                orc(b64["betas"], b64["pose_aa"][:, 3:], b64["pose_aa"][:, :3])
                ts.append(_t.perf_counter() - t0)
            ms_cpu = sorted(ts)[len(ts) // 2] * 1e3
            row["cpu_oracle_axis_angle"] = {"ms_per_forward": ms_cpu, "bodies_per_s": 64 / (ms_cpu * 1e-3),
                                            "threads": torch.get_num_threads()}
        res["smpl_b64"] = row
    # ---- configs[3]: MAF sampling of the 431 down-sampled vertices, B = 1024 per GPU, 14 / 28 / 56 maps, NCHW (contract) and
    #      channels_last; fractions of the HBM roofline on algorithmic bytes
    Bm, Nm, Cm = 1024, 431, 256
    from whmr_b200 import ops as _ops
    pts = torch.from_numpy(syn.make_sample_points(Bm, Nm, seed=2, rank=rank)).to(dev)
    rows = {}
    for (H, Wd) in ((14, 14), (28, 28), (56, 56)):
        feat = torch.randn(Bm, Cm, H, Wd, device=dev)
        alg = Bm * (4 * Cm * (min(4 * Nm, H * Wd) + Nm) + 8 * Nm)
        ms_n = timed(lambda: _ops.sample_bilinear(feat, pts, _ops.LAYOUT_NCHW), 20)
        fl = feat.contiguous(memory_format=torch.channels_last)
        ms_c = timed(lambda: _ops.sample_bilinear(fl, pts, _ops.LAYOUT_NCHW), 20)
        # what a caller pays who holds NCHW maps and converts them for the NHWC kernel ("conversion cost stated")
        ms_conv = timed(lambda: feat.contiguous(memory_format=torch.channels_last), 10)
        rows["%dx%d" % (H, Wd)] = {"nchw_ms": ms_n, "nchw_frac": alg / ms_n / 1e6 / peaks["hbm_gbs"],
                                   "channels_last_ms": ms_c, "channels_last_frac": alg / ms_c / 1e6 / peaks["hbm_gbs"],
                                   "nchw_to_channels_last_conversion_ms": ms_conv, "algorithmic_bytes": alg}
        del feat, fl
    torch.cuda.empty_cache()
    res["maf_sampling_1024x431"] = {"workload": "configs[3]: bilinear sampling of 431 points, 256 channels, %d bodies per GPU" % Bm,
                                    "levels": rows, "scaling": "weak", "peak_GBs": peaks["hbm_gbs"]}
    total = 65536
    lo, hi = shard_bounds(total, rank, world)
    b = syn.make_bodies(hi - lo, seed=5, rank=rank)
    betas, rot = torch.from_numpy(b["betas"]).to(dev), torch.from_numpy(b["rotmat"]).to(dev)
    ms = timed(lambda: ev.joints(betas, rot, True), 5)
    bps = total / (ms * 1e-3)
    res["smpl_sweep_65536"] = {"workload": "configs[2]: SMPL + H36M joint regression, 65,536 bodies in total, %d per GPU" % (hi - lo),
                               "ms": ms, "bodies_per_s": bps, "scaling": "strong",
                               "hbm_frac_algorithmic_per_gpu": bps / world * 84172 / 1e9 / peaks["hbm_gbs"],
                               "tensor_frac_bf16x3_per_gpu": bps / world * 2.0 * KPOSE * 3 * V / 1e12 / (peaks["bf16_tflops_sustained"] / 3)}
    del betas, rot
    # the smaller points of the configs[2] sweep (1k .. 16k bodies in total, same sharding)
    sweep = {}
    for tot in (1024, 4096, 16384):
        lo_s, hi_s = shard_bounds(tot, rank, world)
        bs = syn.make_bodies(max(hi_s - lo_s, 1), seed=5, rank=rank)
        be_s, ro_s = torch.from_numpy(bs["betas"]).to(dev), torch.from_numpy(bs["rotmat"]).to(dev)
        ms_s = timed(lambda: ev.joints(be_s, ro_s, True), 10)
        sweep[str(tot)] = {"ms": ms_s, "bodies_per_s": tot / (ms_s * 1e-3), "per_gpu_bodies": hi_s - lo_s}
        del be_s, ro_s
    sweep["65536"] = {"ms": ms, "bodies_per_s": bps, "per_gpu_bodies": hi - lo}
    res["smpl_sweep"] = {"workload": "configs[2]: SMPL + H36M joint regression, total bodies sharded over the ranks", "points": sweep}
    torch.cuda.empty_cache()
    Nf = 35515
    lo, hi = shard_bounds(Nf, rank, world)
    gt, pr = syn.make_bodies(hi - lo, seed=31, rank=rank), syn.make_bodies(hi - lo, seed=32, rank=rank)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    g_pose, g_betas, p_rot, p_betas = T(gt["pose_aa"]), T(gt["betas"]), T(pr["rotmat"]), T(pr["betas"])

    def shard_pass():
        parts = []
        for a in range(0, hi - lo, 4096):
            r = ev(g_pose[a:a + 4096], g_betas[a:a + 4096], p_rot[a:a + 4096], p_betas[a:a + 4096])
            parts.append(torch.stack([r["mpjpe"], r["pa_mpjpe"], r["pve"]], dim=1))
        return torch.cat(parts) if parts else torch.empty(0, 3, device=dev)

    ms_full = timed(lambda: all_gather_rows(shard_pass(), Nf), 5)
    local = shard_pass()
    ms_gather = timed(lambda: all_gather_rows(local, Nf), 20) if world > 1 else 0.0
    full = all_gather_rows(local, Nf)
    res["eval_pass_35515"] = {"workload": "configs[4]: 35,515 frames, 2 SMPL forwards + H36M read-outs + MPJPE/PA-MPJPE/PVE per frame, "
                                          "%d frames per GPU, all_gather_into_tensor of the [n,3] errors" % (hi - lo),
                              "ms_per_pass": ms_full, "frames_per_s": Nf / (ms_full * 1e-3), "ms_gather_alone": ms_gather,
                              "gathered_shape": list(full.shape), "scaling": "strong",
                              "mean_mm": {"mpjpe": float(full[:, 0].mean()) * 1e3, "pa_mpjpe": float(full[:, 1].mean()) * 1e3,
                                          "pve": float(full[:, 2].mean()) * 1e3}}
    return res


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it may run on."""
    import torch
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_loop_baseline(backbone, n_bodies, reps=3, min_seconds=10.0, max_seconds=30.0):
    import torch
    use_all_host_threads()
    import whmr_b200.synthetic as syn
    from oracle.loop_oracle import LoopOracle, make_cpu_inputs
    model = syn.make_smpl_model(seed=0, weights="random")
    orc = LoopOracle(model, backbone)
    f, p, bb = make_cpu_inputs(n_bodies, backbone)
    orc.step(f, p, bb)
    ts = []
    t_begin = time.perf_counter()
    while True:   # a bounded sample: >= `reps` passes and ~min_seconds of CPU work, never more than max_seconds
        t0 = time.perf_counter()
        orc.step(f, p, bb)
        t1 = time.perf_counter()
        ts.append(t1 - t0)
        spent = t1 - t_begin
        if (len(ts) >= reps and spent >= min_seconds) or spent >= max_seconds:
            break
    reps = len(ts)
    t = sorted(ts)[len(ts) // 2]
    return {"value": n_bodies / t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d bodies of the same loop workload (same generators), median of %d passes, %.2f s/pass; "
                      "oracle = CPU restatement of the reference path (dense SMPL + dense Dmap matmuls + grid_sample), "
                      "torch %s CPU kernels, os.cpu_count()=%d" % (n_bodies, reps, t, torch.__version__, os.cpu_count())}


def torch_gpu_eager_baseline(model, backbone, feats, params, bbox, B, reps=10):
    """The reference's way of computing the path -- dense smplx-style SMPL, dense Dmap / regressor matmuls, F.grid_sample,
    ~100+ ATen launches per SMPL call -- as eager PyTorch on the same GPU and the same device-resident inputs (the oracle
    restatement with its tensors on the device; fp32 matmuls, torch defaults).  A reported baseline like cpu_baseline."""
    import torch
    from oracle.loop_oracle import LoopOracle
    orc = LoopOracle(model, backbone, device="cuda")
    for _ in range(2):
        orc.step(feats, params, bbox)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        orc.step(feats, params, bbox)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": reps,
            "sample": "the full B=%d loop step, eager PyTorch %s on cuda:0 (oracle restatement of the reference path, "
                      "dense matmuls, allow_tf32=%s)" % (B, torch.__version__, torch.backends.cuda.matmul.allow_tf32)}


def whole_loop_leg(model, backbone, dev, B, reps=5):
    """the hot path's share of the WHOLE regressor loop, old vs new.  Around the path stand the
    reference's unchanged PyTorch modules, random-init, eval mode, eager: the deconv stack that turns the backbone output into
    the three feature levels (models/whmr.py:459-500, 560-564) and the three Regressor MLP heads (fc1 / fc2 / decpose /
    decshape / deccam, models/whmr.py:46-55, 117-126).  The path between them -- MAF sampling + reduce_dim, gram-schmidt,
    SMPL, read-outs, weak + full projection, global SMPL -- runs "old" as the oracle restatement of the reference's dense
    eager-PyTorch code on this GPU and "new" through this repo's drop-in modules.  Rank 0, N = 1."""
    import torch
    import torch.nn as nn
    import whmr_b200.synthetic as syn
    from oracle import geometry_oracle as G
    from oracle.loop_oracle import LoopOracle
    from oracle.sampling_oracle import grid_sample_points, reduce_dim
    from whmr_b200.loop import RegressorLoop
    from whmr_b200.maf_extractor import MAF_Extractor
    torch.manual_seed(0)
    vit = backbone == "vitpose"
    cin, (h0, w0) = (768, (16, 12)) if vit else (2048, (7, 7))
    blocks = nn.ModuleList()
    for i in range(3):     # ConvTranspose2d(k=4, s=2, p=1, no bias) + BatchNorm + ReLU, 256 filters each
        blocks.append(nn.Sequential(nn.ConvTranspose2d(cin if i == 0 else 256, 256, 4, 2, 1, bias=False),
                                    nn.BatchNorm2d(256), nn.ReLU(inplace=True)))
    blocks = blocks.to(dev).eval()
    exts = [MAF_Extractor(mesh_downsampling=None).to(dev).eval() for _ in range(3)]
    n_grid = 63 if vit else 64
    heads = nn.ModuleList()
    for i in range(3):
        fd = (n_grid if i == 0 else 67) * 32
        m = nn.ModuleDict({"fc1": nn.Linear(fd + 216 + 13 + 5, 1024), "fc2": nn.Linear(1024, 1024),
                           "decpose": nn.Linear(1024, 216), "decshape": nn.Linear(1024, 10), "deccam": nn.Linear(1024, 3)})
        for k in ("decpose", "decshape", "deccam"):
            nn.init.xavier_uniform_(m[k].weight, gain=0.01)
        heads.append(m)
    heads = heads.to(dev).eval()
    b = syn.make_bodies(B, seed=21)
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    init = {"rotmat": T(b["rotmat"]), "betas": T(b["betas"]), "cam": T(b["cam"])}
    bbox = {k: T(b[k]) for k in ("bbox_height", "center", "orig_shape", "Tz")}
    bbox_info = torch.randn(B, 5, device=dev)
    s_feat = torch.randn(B, cin, h0, w0, device=dev)
    loop = RegressorLoop(model, dev, backbone=backbone)
    orc = LoopOracle(model, backbone, device="cuda")
    convs = [[(c.weight.detach(), c.bias.detach()) for c in e.filters] for e in exts]
    grid = loop.grid
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def mlp(i, feat, st):
        m = heads[i]
        xc = torch.cat([feat, bbox_info, st["rotmat"].reshape(B, -1), st["betas"], st["cam"]], 1)
        xc = m["fc2"](m["fc1"](xc))
        return {"rotmat": (m["decpose"](xc) + st["rotmat"].reshape(B, -1)).view(B, 24, 3, 3),
                "betas": m["decshape"](xc) + st["betas"], "cam": m["deccam"](xc) + st["cam"]}

    def run(new, acc):
        """one pass; acc[k] += ms of section k (deconv / mlp / hot)"""
        marks = []

        def mark(k):
            e = ev(); e.record(); marks.append((k, e))
        torch.cuda.synchronize()     # sections are attributed cleanly: no backlog of the previous pass in the first one
        mark(None)
        feats, x = [], s_feat
        for blk in blocks:
            x = blk(x)
            feats.append(x)
        mark("deconv")
        st = dict(init)
        if new:
            out = loop.head(st["rotmat"], st["betas"], st["cam"], J_regressor=True)
        else:
            out = orc.regressor_outputs(st)
        for it in range(3):
            if new:
                e = exts[it]
                if it == 0:
                    feat = e.sampling(grid, im_feat=feats[0])[0]
                else:
                    e.im_feat, e.cam = feats[it], st["cam"]
                    feat = e(out["markers"], None, None, None, None)[0]
            else:
                pts = grid.unsqueeze(0).expand(B, -1, -1) if it == 0 else G.projection(out["markers"], st["cam"])
                feat = reduce_dim(grid_sample_points(feats[it], pts), convs[it])
            mark("hot")
            st = mlp(it, feat, st)
            mark("mlp")
            if new:
                out = loop.head(st["rotmat"], st["betas"], st["cam"], bbox["bbox_height"], bbox["center"], bbox["orig_shape"],
                                bbox["Tz"], J_regressor=True)
            else:
                out = orc.regressor_outputs(st, bbox)
        if new:      # global call (models/whmr.py:628-651): SMPL + H36M joints of [global_rotmat | body rotations]
            g = loop.head(out["rotmat"], st["betas"], st["cam"], J_regressor=True)
        else:
            g = orc.regressor_outputs({"rotmat": out["rotmat"], "betas": st["betas"], "cam": st["cam"]})
        mark("hot")
        torch.cuda.synchronize()
        for (_, a), (k, e) in zip(marks[:-1], marks[1:]):
            acc[k] = acc.get(k, 0.0) + a.elapsed_time(e)
        return out, g

    res = {}
    with torch.no_grad():
        outs = {}
        for new in (False, True):
            for _ in range(2):
                run(new, {})
            acc = {}
            a, z = ev(), ev()
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                outs[new] = run(new, acc)
            z.record()
            torch.cuda.synchronize()
            tag = "new" if new else "old"
            res["ms_whole_loop_" + tag] = a.elapsed_time(z) / reps
            for k, v in acc.items():
                res["ms_%s_%s" % (k, tag)] = v / reps
        err = float((outs[True][0]["verts"] - outs[False][0]["verts"]).abs().max())
    res["hot_path_share_old"] = res["ms_hot_old"] / res["ms_whole_loop_old"]
    res["hot_path_share_new"] = res["ms_hot_new"] / res["ms_whole_loop_new"]
    res["whole_loop_speedup"] = res["ms_whole_loop_old"] / res["ms_whole_loop_new"]
    res["hot_path_speedup"] = res["ms_hot_old"] / res["ms_hot_new"]
    res["verts_new_vs_old_m"] = err
    res["batch"] = B
    res["note"] = ("eager on both sides (host-launched); deconv stack %d->256->256->256 (ConvTranspose2d k4 s2 + BN + ReLU) from a "
                   "[B,%d,%d,%d] backbone output and the three Regressor MLP heads are plain PyTorch in both arms; 'hot' = MAF "
                   "sampling + reduce_dim, gram-schmidt, SMPL, read-outs, projections, global SMPL: 'old' = the oracle restatement "
                   "of the reference's dense PyTorch code on this GPU, 'new' = this repo's drop-in modules; the ViT / ResNet "
                   "encoder itself is not included" % (cin, cin, h0, w0))
    return res


def run_extra(args):
    """Non-default workloads = the other BASELINE.json configs (parity-test cases; measured here for the record):
      smpl_sweep   configs[2]: SMPL + H36M joint regression, 1k..64k bodies in total, sharded over the ranks
      maf_sampling configs[3]: MAF_Extractor sampling of 431 points, 14/28/56 maps, 1024 bodies per rank, NCHW and NHWC
      eval_pass    configs[4]: 35,515 frames GT SMPL + predicted SMPL + H36M 17->14 + MPJPE/PA-MPJPE/PVE, gathered."""
    import torch
    import torch.distributed as dist
    import whmr_b200.synthetic as syn
    from whmr_b200 import ops
    from whmr_b200.dist import shard_bounds
    from whmr_b200.evaluate import EvalPass
    from whmr_b200.smpl import SMPL
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    model = syn.make_smpl_model(seed=0, weights="random")
    smpl = SMPL(model=model).to(dev)
    res = {"workload": args.workload, "n_gpus": world, "data": "synthetic", "dtype": "f32"}
    if args.workload == "smpl_sweep":
        ev = EvalPass(smpl, model["J_regressor_h36m"])
        rows = {}
        for total in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
            lo, hi = shard_bounds(total, rank, world)
            b = syn.make_bodies(hi - lo, seed=5, rank=rank)
            betas, rot = torch.from_numpy(b["betas"]).to(dev), torch.from_numpy(b["rotmat"]).to(dev)
            ms = timed(lambda: ev.joints(betas, rot, True), 10)
            bps = total / (ms * 1e-3)
            rows[str(total)] = {"ms": ms, "bodies_per_s": bps, "per_gpu_bodies": hi - lo,
                                "hbm_frac_algorithmic": bps / world * 84172 / 1e9 / peaks["hbm_gbs"],
                                "tensor_frac_bf16x3": bps / world * 2.0 * KPOSE * 3 * V / 1e12 / (peaks["bf16_tflops_sustained"] / 3)}
            del betas, rot
        res.update(metric="smpl_h36m_bodies_per_sec", unit="bodies/s", sweep=rows)
    elif args.workload == "maf_sampling":
        B, N, C = 1024, 431, 256
        pts = torch.from_numpy(syn.make_sample_points(B, N, seed=2, rank=rank)).to(dev)
        rows = {}
        from whmr_b200.maf_extractor import MAF_Extractor
        torch.manual_seed(1234)
        ext = MAF_Extractor(mesh_downsampling=None).to(dev).eval()
        mlp_flops = 2.0 * (256 * 128 + 384 * 64 + 320 * 32) * B * N       # algorithmic (fp32) flops of reduce_dim per level
        for (H, W) in ((14, 14), (28, 28), (56, 56)):
            feat = torch.randn(B, C, H, W, device=dev)
            alg = B * (4 * C * (min(4 * N, H * W) + N) + 8 * N)
            ms = timed(lambda: ops.sample_bilinear(feat, pts, ops.LAYOUT_NCHW), 20)
            fl = feat.contiguous(memory_format=torch.channels_last)
            ms2 = timed(lambda: ops.sample_bilinear(fl, pts, ops.LAYOUT_NCHW), 20)   # channels_last -> NHWC kernel, no copy
            rows["%dx%d" % (H, W)] = {"nchw_ms": ms, "nchw_GBs_algorithmic": alg / ms / 1e6, "nchw_frac": alg / ms / 1e6 / peaks["hbm_gbs"],
                                      "channels_last_ms": ms2, "channels_last_GBs_algorithmic": alg / ms2 / 1e6,
                                      "channels_last_frac": alg / ms2 / 1e6 / peaks["hbm_gbs"],
                                      "bodies_per_s_nchw": world * B / (ms * 1e-3)}
            # MAF_Extractor.sampling incl. the reduce_dim MLP: fused kernel vs sampling kernel + PyTorch MLP
            with torch.no_grad():
                row = rows["%dx%d" % (H, W)]
                for tag, f_in in (("nchw", feat), ("channels_last", fl)):
                    ext.fused, ext.return_point_feat = True, False
                    t_f = timed(lambda: ext.sampling(pts, im_feat=f_in), 20)
                    ext.fused, ext.return_point_feat = False, True
                    t_u = timed(lambda: ext.sampling(pts, im_feat=f_in), 10)
                    row["with_reduce_dim_%s" % tag] = {
                        "fused_ms": t_f, "two_step_ms": t_u, "fused_hbm_frac_algorithmic": alg / t_f / 1e6 / peaks["hbm_gbs"],
                        "fused_tensor_frac_3xtf32": mlp_flops / (t_f * 1e-3) / 1e12 / (peaks["bf16_tflops_sustained"] / 2 / 3)}
            del feat, fl
        res.update(metric="maf_sampling_431pts", unit="ms per level (1024 bodies per GPU)", levels=rows)
    elif args.workload == "eval_pass":
        Nf = 35515
        ev = EvalPass(smpl, model["J_regressor_h36m"])
        lo, hi = shard_bounds(Nf, rank, world)
        gt, pr = syn.make_bodies(hi - lo, seed=31, rank=rank), syn.make_bodies(hi - lo, seed=32, rank=rank)
        # per-rank shard resident on the device (the reference's loader feeds batches of 32, evaluate/eval.py:155)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        import numpy as np
        g_pose, g_betas, p_rot, p_betas = T(gt["pose_aa"]), T(gt["betas"]), T(pr["rotmat"]), T(pr["betas"])
        from whmr_b200.dist import all_gather_rows

        def one_pass():
            parts = []
            for a in range(0, hi - lo, 4096):
                r = ev(g_pose[a:a + 4096], g_betas[a:a + 4096], p_rot[a:a + 4096], p_betas[a:a + 4096])
                parts.append(torch.stack([r["mpjpe"], r["pa_mpjpe"], r["pve"]], dim=1))
            return all_gather_rows(torch.cat(parts), Nf)
        ms = timed(one_pass, 5)
        full = one_pass()
        res.update(metric="eval_frames_per_sec", unit="frames/s", value=Nf / (ms * 1e-3), ms_per_pass=ms, frames=Nf,
                   gathered_shape=list(full.shape),
                   mean_mm={"mpjpe": float(full[:, 0].mean()) * 1e3, "pa_mpjpe": float(full[:, 1].mean()) * 1e3,
                            "pve": float(full[:, 2].mean()) * 1e3},
                   note="2 SMPL forwards + H36M read-outs + MPJPE/PA-MPJPE/PVE per frame, NCCL all_gather of [n,3] errors")
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation of the path = the oracle port (smplx / pare / the SMPL
    weights are not installable offline, so the reference itself cannot run)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import whmr_b200.synthetic as syn
    from oracle.loop_oracle import LoopOracle, make_cpu_inputs
    use_all_host_threads()
    K, W = args.steps if args.steps_given else 20, max(1, min(args.warmup, 10))
    # exactly K timed steps of the SAME batch as the GPU arm (256 bodies) while K steps stay within ~3 minutes
    # (~450 bodies/s on 16 threads: K <= ~300); beyond that the per-step sample shrinks
    n = max(8, min(args.batch, args.cpu_sample or args.batch, int(180.0 * 450.0 / max(K, 1))))
    model = syn.make_smpl_model(seed=0, weights="random")
    orc = LoopOracle(model, args.backbone)
    f, p, bb = make_cpu_inputs(n, args.backbone)
    for _ in range(W):
        orc.step(f, p, bb)
    t0 = time.perf_counter()
    for _ in range(K):
        orc.step(f, p, bb)
    t = (time.perf_counter() - t0) / K
    v = n / t
    sample = ("each step = one loop pass over a %d-body sample of the B=256 workload on the host CPU "
              "(torch %s, %d threads)" % (n, torch.__version__, torch.get_num_threads()))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": K, "warmup": W, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "whmr_regressor_loop_B256 (BASELINE configs[1]), CPU arm, %d bodies per step" % n,
                       "batch_per_gpu": args.batch, "parallelism": "cpu"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--backbone", default="vitpose", choices=["vitpose", "res50"])
    ap.add_argument("--gemm-mode", default=None, choices=[None, "fp32_simt", "bf16x3", "3xtf32"])
    ap.add_argument("--cpu-sample", type=int, default=None,
                    help="bodies per CPU pass (default: 64 for the cpu_baseline leg, the full batch for --impl reference)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-depth", type=int, default=3, help="buffer sets of the end-to-end pipeline (>= 2)")
    ap.add_argument("--e2e-compute-streams", type=int, default=1,
                    help="compute streams (each with its own loop object) of the e2e_host_gather leg")
    ap.add_argument("--e2e-host-levels", default="1,2",
                    help="feature levels left in pinned host memory and gathered in place by the sampling kernel in the "
                         "e2e_host_gather leg (comma separated; empty: leg off)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-whole-loop", action="store_true", help="skip the whole-loop (deconv stack + MLP heads around the path) leg")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-eager", action="store_true", help="skip the eager-PyTorch-on-GPU baseline leg")
    ap.add_argument("--skip-sweep", action="store_true")
    ap.add_argument("--skip-other", action="store_true", help="skip the configs[2] / configs[4] legs")
    ap.add_argument("--skip-channels-last", action="store_true")
    ap.add_argument("--skip-train", action="store_true", help="skip the training-step (forward + backward) leg")
    ap.add_argument("--skip-reduce-dim", action="store_true", help="skip the leg with the extractors' MLP fused into sampling")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--workload", default="regressor_loop", choices=["regressor_loop", "smpl_sweep", "maf_sampling", "eval_pass"],
                    help="regressor_loop = BASELINE configs[1] (the bench contract); the others are the remaining configs")
    ap.add_argument("--channels-last", action="store_true",
                    help="feature maps in torch.channels_last memory format (reported separately; the contract layout is NCHW)")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 1000
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "regressor_loop":
        run_extra(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
