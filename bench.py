#!/usr/bin/env python3
"""bench.py -- voxel remesh (mesh -> narrow-band SDF -> marching cubes) throughput on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 5] [--scale 1.0] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one synthetic mesh (BASELINE.json config 5 by default: the ~10M-triangle
noise-displaced UV sphere at 2048^3): MeshToVolume::convert then MarchingCubesMesher::mesh.
`value` = active voxels of the whole job / step time with the triangles already resident in HBM and the output
left in HBM; `e2e` = the same through the public host-buffer API (pinned host triangles in, host vertices out,
copies inside the timed region). With N > 1 the bricks are sharded by contiguous slabs of the sorted brick list
(mesh replicated, no data-path collective except the final all-gather of the triangle buffers).
--impl reference times the CPU restatement of the reference (oracle/, "port": no Rust toolchain here to build the
reference itself) with all host threads on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxel remesh throughput (mesh->SDF + marching cubes, active voxels per second)"
UNIT = "voxels/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0, help="resolution scale of the config (1.0 = BASELINE.json size)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-scale", type=float, default=None,
                    help="scale of the bounded CPU sample (default: the largest scale of CPU_LADDER that fits the time budget on this host, "
                         "extrapolated from two measured calibration passes: ~25 s for the cpu_baseline pass of our arm, ~150 s for the reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sign-propagation", action="store_true",
                    help="A/B switch: per-voxel winding numbers even on closed meshes (BS_FLAG_SIGN_PROPAGATION = 0); recorded in config")
    ap.add_argument("--gather", default="fused", choices=["fused", "p2p", "nccl"],
                    help="N > 1 output exchange: fused = the marching-cubes emit kernel stores every triangle into all ranks' buffers over NVLink "
                         "(bs_mesh_mc_count + bs_mesh_mc_emit_push, CUDA IPC mapped peer memory: the exchange rides on the emission); p2p = extraction, "
                         "then one kernel copies the rank's slice into all peers' buffers (bs_context_push_out_verts); nccl = torch.distributed all-gather (the baseline)")
    ap.add_argument("--remesh-slabs", type=int, default=0, help="N = 1 e2e leg: slabs of the pipelined remesh call (0 = the library's choice)")
    ap.add_argument("--e2e-two-calls", action="store_true", help="N = 1 e2e leg through bs_mesh_to_volume + bs_mesh_mc_device + bs_context_copy_out_verts (no overlap) instead of bs_voxel_remesh_into")
    ap.add_argument("--io", action="store_true", help="time the rows either side of the path on the --config mesh: STL decode / encode, merge_points, ActiveVoxelsMesher")
    ap.add_argument("--ops", action="store_true", help="time the CSG (config 2) / offset (3) / dual contouring (4) rows instead of the remesh")
    return ap.parse_args()


def workload(cfg, scale):
    from baby_shark_b200 import synth
    if cfg not in (1, 3, 4, 5):
        raise SystemExit("the remesh bench takes --config 1, 3, 4 or 5")
    tris, vs, desc = synth.config_mesh(cfg, scale)
    return np.ascontiguousarray(tris, np.float32), float(vs), desc


def run_ops(args):
    """--ops: the other rows of the path on their BASELINE.json configs, single GPU, volumes resident in HBM:
    config 2 = union + subtract of two tori then MC; config 3 = offset(+2 voxels) and offset(-2 voxels) then MC;
    config 4 = convert + dual contouring. One JSON line with per-stage device times and rooflines."""
    import torch
    import baby_shark_b200 as B
    from baby_shark_b200 import synth
    L = B.load_library()
    ctx = B.Context(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    cfg = args.config
    mesh, vs, desc = synth.config_mesh(cfg, args.scale)
    conv = lambda t: B.MeshToVolume(ctx).with_voxel_size(vs).convert(t)  # noqa: E731
    stages, work, rl = {}, {}, []

    def timed(fn, reps):
        """average per-stage device ms over `reps` runs of fn (fn returns the stats dict of the call of interest)"""
        acc = {}
        for _ in range(reps):
            for k, v in fn().items():
                acc[k] = acc.get(k, 0.0) + v / reps
        return acc

    mc = B.MarchingCubesMesher().with_voxel_size(vs)
    if cfg == 2:
        a, b = conv(mesh[0]), conv(mesh[1])
        for op in ("union", "subtract"):
            for _ in range(args.warmup):
                getattr(a.clone(), op)(b.clone())
            def one(op=op):
                r = getattr(a.clone(), op)(b.clone())
                st = ctx.last_stats()
                return st
            st = timed(one, args.steps)
            stages.update({op + "_" + k: v for k, v in st.items() if k.endswith("_ms")})
            work.update({op + "_" + k: v for k, v in st.items() if not k.endswith("_ms")})
            by = 6336.0 * st["n_merged_bricks"] + 4224.0 * (st["n_out_bricks"] - st["n_merged_bricks"])
            rl.append({"kernel": "k_csg_bricks (%s)" % op, "bound": "hbm", "achieved": by / (st["csg_bricks_ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "ms": st["csg_bricks_ms"], "algorithmic": "6336 B x merged + 4224 B x copied bricks", "traffic": None})
        metric, unit = "CSG union+subtract of two tori (flood fill + directory + brick merge)", "ms"
        value = stages["union_total_ms"] + stages["subtract_total_ms"]
    elif cfg == 3:
        a = conv(mesh)
        for sgn, name in ((2.0, "offset_plus"), (-2.0, "offset_minus")):
            for _ in range(args.warmup):
                a.clone().offset(sgn * vs)
            def one(sgn=sgn):
                a.clone().offset(sgn * vs)
                return ctx.last_stats()
            st = timed(one, args.steps)
            stages.update({name + "_" + k: v for k, v in st.items() if k.endswith("_ms")})
            work.update({name + "_" + k: v for k, v in st.items() if not k.endswith("_ms")})
            by = 8.0 * st["n_out_bricks"] * 4224.0
            rl.append({"kernel": "k_sweep x %d launches (%s)" % (st["n_sweep_launches"], name), "bound": "hbm", "achieved": by / (st["offset_sweep_ms"] * 1e-3) / 1e9, "peak": hbm_peak,
                       "unit": "GB/s", "ms": st["offset_sweep_ms"], "algorithmic": "8 sweeps x n_out_bricks x 4224 B (dependency-bound: expected far below the roofline)", "traffic": None})
        metric, unit = "offset +2 and -2 voxels of a ~1M-triangle sphere at 1024^3 (prune + 8 sweeps + shift)", "ms"
        value = stages["offset_plus_total_ms"] + stages["offset_minus_total_ms"]
    else:
        a = conv(mesh)
        dc = B.DualContouringMesher().with_voxel_size(vs)
        for _ in range(args.warmup):
            dc.mesh(a)
        def one():
            dc.mesh(a)
            return ctx.last_stats()
        st = timed(one, args.steps)
        stages.update({k: v for k, v in st.items() if k.endswith("_ms")})
        work.update({k: v for k, v in st.items() if not k.endswith("_ms")})
        by = st["n_bricks"] * (11 ** 3 * 4 + 64) + 36.0 * st["n_out_tris"]
        t = st["dc_cells_ms"] + st["dc_count_ms"] + st["dc_emit_ms"]
        rl.append({"kernel": "k_dc_cells + k_dc_quads", "bound": "hbm", "achieved": by / (t * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "ms": t,
                   "algorithmic": "n_bricks x 5388 B + 36 B x n_out_tris", "traffic": None})
        metric, unit = "dual contouring of a 2M-triangle noise sphere at 1024^3 (triangles per second)", "tris/s"
        value = st["n_out_tris"] / (t * 1e-3)
    for r in rl:
        r["frac"] = r["achieved"] / r["peak"]
    print(json.dumps({"metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": unit != "ms",
                      "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": "config %d: %s, scale %g" % (cfg, desc, args.scale), "voxel_size": vs},
                      "stage_ms": stages, "work": work, "rooflines": rl, "roofline": max(rl, key=lambda r: r["ms"]) if rl else None}))


def run_io(args):
    """--io: STL bytes -> device triangles -> volume -> MC soup -> {STL bytes, merged vertices}, plus ActiveVoxelsMesher;
    everything resident in HBM, per-kernel device times and HBM rooflines in one JSON line."""
    import torch
    import baby_shark_b200 as B
    L = B.load_library()
    ctx = B.Context(0)
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0
    tris, vs, desc = workload(args.config, args.scale)
    n = tris.shape[0]
    rec = np.zeros((n, 50), np.uint8)
    rec[:, 12:48] = tris.view(np.uint8).reshape(n, 36)
    stl = torch.from_numpy(np.concatenate([np.zeros(80, np.uint8), np.frombuffer(np.uint32(n).tobytes(), np.uint8), rec.reshape(-1)])).cuda()
    del rec
    stages, rl = {}, []

    def avg(fn, key):
        for _ in range(args.warmup):
            fn()
        acc = 0.0
        for _ in range(args.steps):
            fn()
            acc += ctx.last_stats()[key] / args.steps
        return acc

    keep = {}

    def decode():
        if keep.get("tris"):
            L.bs_device_free(ctx._h, keep["tris"])
        p, m = C.c_void_p(), C.c_size_t()
        ctx.check(L.bs_stl_decode_device(ctx._h, C.c_void_p(stl.data_ptr()), stl.numel(), C.byref(p), C.byref(m)))
        keep["tris"], keep["n"] = p, m.value
    stages["stl_decode_ms"] = avg(decode, "stl_decode_ms")
    rl.append({"kernel": "k_stl_decode", "ms": stages["stl_decode_ms"], "algorithmic": "86 B x %d triangles" % n, "bytes": 86.0 * n})
    h = C.c_void_p()
    ctx.check(L.bs_mesh_to_volume_device(ctx._h, keep["tris"], n, vs, 0, C.byref(h)))
    dv, nv = C.c_void_p(), C.c_size_t()
    ctx.check(L.bs_mesh_mc_device(h, vs, C.byref(dv), C.byref(nv)))
    n_out = nv.value // 3

    def encode():
        p, m = C.c_void_p(), C.c_size_t()
        ctx.check(L.bs_stl_encode_device(ctx._h, dv, nv.value, C.byref(p), C.byref(m)))
        L.bs_device_free(ctx._h, p)
    stages["stl_encode_ms"] = avg(encode, "stl_encode_ms")
    rl.append({"kernel": "k_stl_encode", "ms": stages["stl_encode_ms"], "algorithmic": "86 B x %d triangles" % n_out, "bytes": 86.0 * n_out})
    uniq = {}

    def merge():
        pu, pi, nu = C.c_void_p(), C.c_void_p(), C.c_size_t()
        ctx.check(L.bs_merge_points_device(ctx._h, dv, nv.value, C.byref(pu), C.byref(nu), C.byref(pi)))
        uniq["n"] = nu.value
        L.bs_device_free(ctx._h, pu)
        L.bs_device_free(ctx._h, pi)
    stages["merge_points_ms"] = avg(merge, "merge_points_ms")
    rl.append({"kernel": "k_mp_insert + k_mp_first + scan + k_mp_emit", "ms": stages["merge_points_ms"],
               "algorithmic": "16 B x %d points + 12 B x %d unique" % (nv.value, uniq["n"]), "bytes": 16.0 * nv.value + 12.0 * uniq["n"]})
    faces = {}

    def boxes():
        p, m = C.c_void_p(), C.c_size_t()
        ctx.check(L.bs_mesh_active_voxels_device(h, C.byref(p), C.byref(m)))
        faces["n"] = m.value // 6
        L.bs_device_free(ctx._h, p)
    stages["active_voxels_ms"] = avg(boxes, "active_voxels_ms")
    nb = ctx.last_stats().get("n_bricks", 0.0)
    rl.append({"kernel": "k_active_voxels<count> + scan + k_active_voxels<emit>", "ms": stages["active_voxels_ms"],
               "algorithmic": "72 B x %d exposed faces (+ 7 x 64 B of masks per brick, twice)" % faces["n"], "bytes": 72.0 * faces["n"]})
    L.bs_volume_free(h)
    for r in rl:
        r.update({"bound": "hbm", "achieved": r.pop("bytes") / (r["ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "traffic": None})
        r["frac"] = r["achieved"] / r["peak"]
    print(json.dumps({"metric": "STL decode + encode + merge_points + ActiveVoxelsMesher on the --config mesh", "value": sum(stages.values()), "unit": "ms", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u8/f32/u32", "data": "synthetic",
                      "config": {"workload": "config %d: %s, scale %g" % (args.config, desc, args.scale), "n_triangles": int(n), "n_out_triangles": int(n_out), "n_unique_vertices": uniq["n"]},
                      "stage_ms": stages, "rooflines": rl, "roofline": max(rl, key=lambda r: r["ms"])}))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc, self.t_begin = index, [], False, None, None

    def mark_begin(self):
        """nvidia-smi is started before the warm-up (its start-up takes driver locks that stall launches for 100s of ms);
        only samples taken after this call -- the timed region -- are kept."""
        self.t_begin = time.time()

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                if self.t_begin is not None:  # rows before mark_begin() belong to the warm-up
                    self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port(tris, vs, threads):
    """One pass of the reference's CPU path (oracle port): convert + MC. Returns (active voxels, triangles, seconds, stats)."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    vol, st = O.mesh_to_volume(tris, vs, 0, threads)
    t1 = time.perf_counter()
    verts = O.marching_cubes(vol, vs)
    dt = time.perf_counter() - t0
    st = dict(st)
    st["t_mc"] = time.perf_counter() - t1
    return st["n_active"], verts.shape[0] // 3, dt, st


CPU_LADDER = (1.0, 0.75, 0.5, 0.375, 0.25, 0.1875, 0.125, 0.09375, 0.0625)


def cpu_pick_scale(cfg, threads, n_pass, budget_s):
    """Largest scale of CPU_LADDER whose n_pass passes fit in budget_s, from two MEASURED calibration passes (scales 0.0625 and
    0.125): seconds(scale) = t_0.125 * (scale / 0.125) ** p with the measured exponent p. Returns (scale, p, calibration dict)."""
    import math
    t = {}
    for sc in (0.0625, 0.125):
        tris, vs, _ = workload(cfg, sc)
        t[sc] = cpu_port(tris, vs, threads)[2]
    p = max(2.0, min(3.5, math.log(t[0.125] / t[0.0625]) / math.log(2.0)))
    budget = budget_s - t[0.0625] - t[0.125]
    scale = next((sc for sc in CPU_LADDER if t[0.125] * (sc / 0.125) ** p * n_pass <= budget), 0.0625)
    return scale, p, {"seconds_at_0.0625": round(t[0.0625], 2), "seconds_at_0.125": round(t[0.125], 2)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_warm = max(0, min(args.warmup, 1))
    n_pass = max(1, args.steps) + n_warm
    calib, expo = None, None
    if args.cpu_scale is None:  # the largest sample whose steps + warm-up fit in ~150 s on THIS host, measured not tabulated
        args.cpu_scale, expo, calib = cpu_pick_scale(args.config, cores, n_pass, 150.0)
    tris, vs, desc = workload(args.config, args.cpu_scale)
    for _ in range(n_warm):
        cpu_port(tris, vs, cores)
    tot_v, tot_t, stage = 0, 0.0, {}
    for _ in range(max(1, args.steps)):
        nv, nt, dt, st = cpu_port(tris, vs, cores)
        tot_v += nv
        tot_t += dt
        for k in ("t_subdivide", "t_tree", "t_udf", "t_sign", "t_mc"):
            stage[k] = stage.get(k, 0.0) + st[k] / max(1, args.steps)
    value = tot_v / tot_t
    sample = "config %d at scale %g: %s (%d triangles, %d active voxels), convert + MC per step" % (args.config, args.cpu_scale, desc, tris.shape[0], nv)
    cb = {"value": value, "unit": UNIT, "cores": cores, "threads": cores, "kind": "port", "sample": sample,
          "stage_seconds": {k: round(v, 3) for k, v in stage.items()},
          "note": "C++ restatement of the reference (no Rust toolchain): distance evaluations and the sign pass run on all host threads, tree build / min-merge / MC are serial as in the reference"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": "config %d: %s" % (args.config, workload_desc(args)), "sample": sample, "sample_scale": args.cpu_scale},
        "same_config": bool(args.cpu_scale == args.scale), "time_exponent_vs_scale": expo, "calibration": calib,
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_desc(args):
    names = {1: "voxel_remeshing example, assets/bunny.stl at voxel_size 0.01", 3: "UV sphere ~1.0M triangles at 1024^3", 4: "noise-displaced UV sphere 2.0M triangles at 1024^3",
             5: "noise-displaced UV sphere ~10.0M triangles at 2048^3"}
    return "%s, voxel remesh (convert + MC33), scale %g" % (names[args.config], args.scale)


def main():
    args = parse()
    if args.no_sign_propagation:
        os.environ["BSHARK_SIGN_PROPAGATION"] = "0"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.ops:
        if rank == 0:
            run_ops(args)
        return
    if args.io:
        if rank == 0:
            run_io(args)
        return

    import torch
    import torch.distributed as dist
    import baby_shark_b200 as B
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = B.load_library()
    ctx = B.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    tris, vs, desc = workload(args.config, args.scale)
    n_tris = tris.shape[0]
    h_tris = torch.from_numpy(tris).pin_memory()                     # pinned host input for the e2e leg
    d_tris = h_tris.to("cuda", non_blocking=False)                   # resident input for the `value` leg
    fp = C.POINTER(C.c_float)

    from baby_shark_b200.shard import all_gather_varlen
    gather_buf = {"out": None, "counts": None}

    class _DeviceF32:
        """zero-copy torch view of a device buffer the library owns (CUDA array interface)"""
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}

    peer = {"cap": 0, "own": None, "ptrs": None, "fence": None}

    def exchange_counts(n_floats):
        cnt = torch.tensor([n_floats], dtype=torch.int64, device="cuda")
        got = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(got, cnt)
        return [int(c.item()) for c in got]

    def peer_buffers(total_floats):
        """(re)allocate this rank's result buffer as an IPC-shareable allocation and map every peer's (bs_ipc_*)"""
        if peer["ptrs"] is not None:
            for r, p in enumerate(peer["ptrs"]):
                if r != rank:
                    ctx.check(L.bs_ipc_close(ctx._h, C.c_void_p(p)))
            dist.barrier()
            ctx.check(L.bs_ipc_free(ctx._h, C.c_void_p(peer["own"])))
        cap = int(total_floats * 1.1) + 1024
        own, handle = C.c_void_p(), C.create_string_buffer(64)
        ctx.check(L.bs_ipc_alloc(ctx._h, cap * 4, C.byref(own), handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw))
        ptrs = []
        for r in range(world):
            if r == rank:
                ptrs.append(own.value)
            else:
                q = C.c_void_p()
                ctx.check(L.bs_ipc_open(ctx._h, handles[r], C.byref(q)))
                ptrs.append(q.value)
        peer.update({"cap": cap, "own": own.value, "ptrs": ptrs, "arr": (C.c_void_p * world)(*ptrs), "fence": torch.zeros(1, device="cuda")})
        dist.barrier()

    def gather(dv_ptr, n_floats):
        """the path's one exchange: every rank ends up with the whole triangle soup, in rank order. The per-rank sizes of a
        steady workload are those of the previous step: no count exchange, no host synchronisation inside the step.
        p2p (default): each rank stores its slice into every rank's result buffer with one kernel (peer memory over NVLink,
        bs_context_push_out_verts) and a one-element all-reduce on the same stream is the "everyone has delivered" fence;
        nccl: torch.distributed all-gather straight out of the library's buffer (the baseline)."""
        counts = gather_buf["counts"]
        if counts is not None and counts[rank] != n_floats:
            counts = None
        if args.gather == "p2p":
            if counts is None:
                counts = exchange_counts(n_floats)
                gather_buf["counts"] = counts
            total = sum(counts)
            if peer["cap"] < total:
                peer_buffers(total)
            ctx.check(L.bs_context_push_out_verts(ctx._h, peer["arr"], world, sum(counts[:rank]), n_floats))
            with torch.cuda.stream(stream):
                dist.all_reduce(peer["fence"])
            gather_buf["out"] = torch.as_tensor(_DeviceF32(peer["own"], total), device="cuda")
            return gather_buf["out"]
        local = torch.as_tensor(_DeviceF32(dv_ptr, n_floats), device="cuda")
        with torch.cuda.stream(stream):
            full, counts = all_gather_varlen(local, out=gather_buf["out"], counts=counts)
        if gather_buf["out"] is None or gather_buf["out"].data_ptr() != full.data_ptr():
            gather_buf["out"] = torch.empty(int(full.numel() * 1.1) + 16, dtype=torch.float32, device="cuda")
        gather_buf["counts"] = counts
        return full

    phase_ev = []  # (convert done, MC done, gather done) events of every timed step, on the library's stream
    mc_extra = {}  # fused exchange: stage times of the count call (the emit call's stats replace them)

    def step_device(record=False):
        """convert + MC (+ output exchange when sharded), input and output resident in HBM; returns local vertex count."""
        h = C.c_void_p()
        ctx.check(L.bs_mesh_to_volume_sharded(ctx._h, C.c_void_p(d_tris.data_ptr()), n_tris, vs, 0, rank, world, C.byref(h)))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if record else None
        if record:
            ev[0].record(stream)
        dv, nv = C.c_void_p(), C.c_size_t()
        if world > 1 and args.gather == "fused":
            # count -> (counts of a steady workload: those of the previous step) -> emit straight into every rank's buffer
            st = L.bs_mesh_mc_count(h, vs, C.byref(nv))
            if st == 0:
                if record:
                    mc_extra["mc_count_ms"] = mc_extra.get("mc_count_ms", 0.0) + ctx.last_stats().get("mc_count_ms", 0.0)
                n_floats = nv.value * 3
                counts = gather_buf["counts"]
                if counts is None or counts[rank] != n_floats:
                    counts = exchange_counts(n_floats)
                    gather_buf["counts"] = counts
                total = sum(counts)
                if peer["cap"] < total:
                    peer_buffers(total)
                if record:
                    ev[1].record(stream)
                st = L.bs_mesh_mc_emit_push(h, peer["arr"], world, sum(counts[:rank]), peer["cap"])
                with torch.cuda.stream(stream):
                    dist.all_reduce(peer["fence"])
                gather_buf["out"] = torch.as_tensor(_DeviceF32(peer["own"], total), device="cuda")
            L.bs_volume_free(h)
            ctx.check(st)
            if record:
                ev[2].record(stream)
                phase_ev.append(ev)
            return None, nv.value, None
        st = L.bs_mesh_mc_device(h, vs, C.byref(dv), C.byref(nv))
        L.bs_volume_free(h)
        ctx.check(st)
        if record:
            ev[1].record(stream)
        if world > 1:
            gather(dv.value, nv.value * 3)
        if record:
            ev[2].record(stream)
            phase_ev.append(ev)
        return dv.value, nv.value, None

    h_out = None
    shm = {"t": None, "offs": None, "path": None}
    d_in = torch.empty_like(d_tris) if world > 1 else None

    def shared_host_buffer(n_floats_total):
        """N > 1: one page-locked host buffer shared by all ranks (a /dev/shm file mapped and cudaHostRegister-ed by every
        process): each rank copies ITS slice of the result to the host over its own PCIe link, rank 0 reads the whole mesh."""
        import mmap
        path = "/dev/shm/bshark_bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), n_floats_total)
        nbytes = max(4, n_floats_total * 4)
        if rank == 0:
            with open(path, "wb") as f:
                f.truncate(nbytes)
        dist.barrier()
        f = open(path, "r+b")
        mm = mmap.mmap(f.fileno(), nbytes)
        arr = np.frombuffer(mm, dtype=np.float32)
        t = torch.from_numpy(arr)
        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)
        assert int(rc) == 0, "cudaHostRegister failed: %s" % rc
        dist.barrier()
        if rank == 0:
            os.unlink(path)
        return t

    def step_e2e():
        nonlocal h_out
        h = C.c_void_p()
        if world == 1:
            ctx.check(L.bs_mesh_to_volume(ctx._h, C.cast(h_tris.data_ptr(), fp), n_tris, vs, 0, C.byref(h)))
        else:  # replicated mesh: uploaded once (rank 0), broadcast over NVLink, every rank keeps its slab
            with torch.cuda.stream(stream):
                if rank == 0:
                    d_in.copy_(h_tris, non_blocking=True)
                dist.broadcast(d_in, src=0)
            ctx.check(L.bs_mesh_to_volume_sharded(ctx._h, C.c_void_p(d_in.data_ptr()), n_tris, vs, 0, rank, world, C.byref(h)))
        dv, nv = C.c_void_p(), C.c_size_t()
        st = L.bs_mesh_mc_device(h, vs, C.byref(dv), C.byref(nv))
        L.bs_volume_free(h)
        ctx.check(st)
        n_floats = nv.value * 3
        if world > 1:  # every rank writes its slice of the soup into the shared pinned host buffer
            if shm["t"] is None or shm["counts"][rank] != n_floats:
                cnt = torch.tensor([n_floats], dtype=torch.int64, device="cuda")
                got = [torch.zeros_like(cnt) for _ in range(world)]
                dist.all_gather(got, cnt)
                shm["counts"] = [int(c.item()) for c in got]
                shm["offs"] = [sum(shm["counts"][:r]) for r in range(world + 1)]
                shm["t"] = shared_host_buffer(shm["offs"][world])
            ctx.check(L.bs_context_copy_out_verts(ctx._h, C.c_void_p(shm["t"].data_ptr() + 4 * shm["offs"][rank]), n_floats))
            dist.barrier()  # the step is done when every slice has landed
            return nv.value
        if h_out is None or h_out.numel() < n_floats:
            h_out = torch.empty(int(n_floats * 1.1) + 16, dtype=torch.float32).pin_memory()
        ctx.check(L.bs_context_copy_out_verts(ctx._h, C.c_void_p(h_out.data_ptr()), n_floats))
        return nv.value

    def step_e2e_remesh():
        """N = 1: VoxelRemesher::remesh as ONE library call, pinned host triangles in, pinned host vertices out; the library
        converts + extracts slab by slab and overlaps the read-back of one slab with the kernels of the next"""
        nonlocal h_out
        if h_out is None:
            h_out = torch.empty(int(n_verts_local * 3 * 1.1) + 1024, dtype=torch.float32).pin_memory()
        nf = C.c_size_t()
        for _ in range(2):
            st = L.bs_voxel_remesh_into(ctx._h, C.c_void_p(h_tris.data_ptr()), n_tris, vs, 0, args.remesh_slabs, C.c_void_p(h_out.data_ptr()), h_out.numel(), C.byref(nf))
            if st == 3 and nf.value > h_out.numel():
                h_out = torch.empty(int(nf.value * 1.1) + 1024, dtype=torch.float32).pin_memory()
                continue
            break
        ctx.check(st)
        return nf.value // 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one instrumented pass (outside any timed region): work counters for the rooflines + active voxel count
    L.bs_context_set_flag(ctx._h, 1, 1)
    h = C.c_void_p()
    ctx.check(L.bs_mesh_to_volume_sharded(ctx._h, C.c_void_p(d_tris.data_ptr()), n_tris, vs, 0, rank, world, C.byref(h)))
    work = ctx.last_stats()
    L.bs_context_set_flag(ctx._h, 1, 0)
    L.bs_volume_free(h)
    n_active_local = work.get("n_active", 0.0)
    if world > 1:  # sharded volumes carry halo bricks: count the job's active voxels once on the unsharded volume
        h = C.c_void_p()
        ctx.check(L.bs_mesh_to_volume_device(ctx._h, C.c_void_p(d_tris.data_ptr()), n_tris, vs, 0, C.byref(h)))
        na = C.c_size_t()
        ctx.check(L.bs_volume_counts(h, None, C.byref(na), None, None))
        L.bs_volume_free(h)
        n_active_local = float(na.value) / world  # summed over ranks below

    # ---- value leg: K steps, device resident ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    stage_ms = {}
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    n_verts_local = 0
    launches0 = L.bs_kernel_launch_count()
    marks = []
    for _ in range(args.steps):
        _, nv, _ = step_device(record=True)
        n_verts_local = nv
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record(stream)
        for k, v in ctx.last_stats().items():   # MC stage timings of this step (convert's were overwritten; re-read below)
            if k.endswith("_ms") and k != "total_ms":
                stage_ms[k] = stage_ms.get(k, 0.0) + v
    e1.record(stream)
    launches = L.bs_kernel_launch_count() - launches0  # this rank's own kernels launched inside the timed region (library sorts / scans excluded)
    barrier()
    clocks = sampler.finish()
    ms_total = e0.elapsed_time(e1)
    step_ms = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    starts = [e0] + marks[:-1]
    phase_ms = {"convert_call": sum(a.elapsed_time(ev[0]) for a, ev in zip(starts, phase_ev)) / args.steps,
                ("mc_count_call" if (world > 1 and args.gather == "fused") else "mc_call"): sum(ev[0].elapsed_time(ev[1]) for ev in phase_ev) / args.steps,
                ("mc_emit_into_all_ranks" if (world > 1 and args.gather == "fused") else "all_gather"): sum(ev[1].elapsed_time(ev[2]) for ev in phase_ev) / args.steps}

    # per-stage device times of convert (CUDA events on the library's stream), averaged over a few extra passes
    conv_ms = {}
    reps = max(1, min(3, args.steps))
    for _ in range(reps):
        h = C.c_void_p()
        ctx.check(L.bs_mesh_to_volume_sharded(ctx._h, C.c_void_p(d_tris.data_ptr()), n_tris, vs, 0, rank, world, C.byref(h)))
        for k, v in ctx.last_stats().items():
            if k.endswith("_ms") and k != "total_ms":
                conv_ms[k] = conv_ms.get(k, 0.0) + v / reps
        L.bs_volume_free(h)
    mc_ms = {k: v / args.steps for k, v in stage_ms.items()}
    mc_ms.update({k: v / args.steps for k, v in mc_extra.items()})

    # ---- e2e leg ---------------------------------------------------------------------------------------------------
    e2e_step = step_e2e_remesh if (world == 1 and not args.e2e_two_calls) else step_e2e
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nv_e2e = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_stats = ctx.last_stats() if e2e_step is step_e2e_remesh else None

    # ---- self-check: the vertices the e2e leg just delivered to the host (and, on one GPU, the volume itself) must be the
    # CPU oracle's, bit for bit and in order (fingerprints written by tests/golden/make_config_hashes.py at this size) ------
    verify_out = None
    gathered_host = None
    if world > 1:  # one more exchange, run by EVERY rank (it ends in a collective fence); rank 0 checks what it received
        step_device()
        torch.cuda.synchronize()
        if rank == 0:
            gathered_host = gather_buf["out"].cpu()
        barrier()
    if rank == 0:
        from baby_shark_b200 import verify
        try:
            golden = json.load(open(os.path.join(ROOT, "tests", "golden", "config_hashes.json"))).get("cfg%d@%g" % (args.config, args.scale))
        except Exception:
            golden = None
        if golden is not None:
            host = shm["t"][: shm["offs"][world]] if world > 1 else h_out[: int(nv_e2e) * 3]
            got = verify.fingerprint_soup(host.numpy())
            if world > 1 and gathered_host is not None:  # the device-resident result of the value leg's exchange as well
                g2 = verify.fingerprint_soup(gathered_host.numpy())
                got["gathered_ok"] = bool(g2 == {k: got[k] for k in g2})
                golden = dict(golden, gathered_ok=True)
            if world == 1:
                hv = C.c_void_p()
                ctx.check(L.bs_mesh_to_volume_device(ctx._h, C.c_void_p(d_tris.data_ptr()), n_tris, vs, 0, C.byref(hv)))
                got.update(verify.fingerprint_volume(B.Volume(hv, ctx).download()))
            bad = sorted(k for k, v in got.items() if golden.get(k) != v)
            verify_out = {"verified": not bad, "checked": sorted(got), "mismatch": bad, "golden": "tests/golden/config_hashes.json cfg%d@%g (CPU oracle, %s)" % (args.config, args.scale, golden.get("desc", ""))}

    # ---- e2e, indexed output (single GPU): the same remesh handed to an indexed mesh type -- MC + merge_points on the
    # device, unique points + indices read back (half the bytes of the soup, no host hash pass) -------------------------
    e2e_indexed = None
    if world == 1:
        h_pts = h_idx = None

        def step_e2e_indexed():
            nonlocal h_pts, h_idx
            h = C.c_void_p()
            ctx.check(L.bs_mesh_to_volume(ctx._h, C.cast(h_tris.data_ptr(), fp), n_tris, vs, 0, C.byref(h)))
            dp, di, npnt, nidx = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
            st = L.bs_mesh_mc_indexed_device(h, vs, C.byref(dp), C.byref(npnt), C.byref(di), C.byref(nidx))
            L.bs_volume_free(h)
            ctx.check(st)
            if h_pts is None or h_pts.numel() < npnt.value * 3:
                h_pts = torch.empty(int(npnt.value * 3.3) + 16, dtype=torch.float32).pin_memory()
            if h_idx is None or h_idx.numel() < nidx.value:
                h_idx = torch.empty(int(nidx.value * 1.1) + 16, dtype=torch.int32).pin_memory()
            ctx.check(L.bs_copy_to_host(ctx._h, dp, C.c_void_p(h_pts.data_ptr()), npnt.value * 12))
            ctx.check(L.bs_copy_to_host(ctx._h, di, C.c_void_p(h_idx.data_ptr()), nidx.value * 4))
            L.bs_device_free(ctx._h, dp)
            L.bs_device_free(ctx._h, di)
            return npnt.value, nidx.value
        for _ in range(2):
            step_e2e_indexed()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            n_pts, n_idx = step_e2e_indexed()
        torch.cuda.synchronize()
        ms_i = (time.perf_counter() - t0) * 1e3 / args.steps
        e2e_indexed = {"ms_per_step": ms_i, "h2d_bytes_per_step": int(tris.nbytes), "d2h_bytes_per_step": int(n_pts * 12 + n_idx * 4),
                       "n_unique_vertices": int(n_pts), "what": "bs_mesh_to_volume + bs_mesh_mc_indexed_device + read-back of unique points and indices"}

    # ---- reduce over ranks ---------------------------------------------------------------------------------------
    t = torch.tensor([ms_total, e2e_s * 1e3, n_active_local, float(n_verts_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, e2e_ms = float(tmax[0]), float(tmax[1])
        n_active, n_verts = float(tsum[2]), float(tsum[3])
    else:
        e2e_ms = e2e_s * 1e3
        n_active, n_verts = n_active_local, float(n_verts_local)

    per_rank = None
    if world > 1:
        per_rank = [None] * world
        mine = dict(conv_ms)
        mine.update({k: v / args.steps for k, v in stage_ms.items()})
        mine.update({k: v for k, v in work.items() if k.startswith("fwn_") or k.startswith("n_")})
        dist.all_gather_object(per_rank, mine)
    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = n_active / (ms_per_step * 1e-3)
        e2e_value = n_active / (e2e_ms / args.steps * 1e-3)
        prop = torch.cuda.get_device_properties(local_rank)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        sm_mhz = clocks.get("sm_max_mhz") or 1965.0
        fp32_peak = prop.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12  # TFLOP/s with FMA at max clock
        rl = []
        if "udf_ms" in conv_ms and work.get("n_eval"):
            fl = 80.0 * work["n_eval"]
            rl.append({"kernel": "k_eval (point-triangle distances, scatter-min)", "bound": "fp32", "achieved": fl / (conv_ms["udf_ms"] * 1e-3) / 1e12,
                       "peak": fp32_peak, "unit": "TFLOP/s", "ms": conv_ms["udf_ms"], "algorithmic": "80 FLOP x n_eval=%d" % work["n_eval"], "traffic": None})
        if "sign_ms" in conv_ms and work.get("fwn_voxels") and not work.get("sign_propagation"):
            fl = 60.0 * work["fwn_far"] + 100.0 * work["fwn_exact_tris"] + 10.0 * work["fwn_visits"]
            rl.append({"kernel": "k_sign (fast winding numbers)", "bound": "fp32", "achieved": fl / (conv_ms["sign_ms"] * 1e-3) / 1e12, "peak": fp32_peak,
                       "unit": "TFLOP/s", "ms": conv_ms["sign_ms"], "traffic": None,
                       "algorithmic": "per voxel: %.1f visits, %.1f far, %.1f exact tris" % tuple(work[k] / work["fwn_voxels"] for k in ("fwn_visits", "fwn_far", "fwn_exact_tris"))})
        if work.get("sign_propagation") and "sign_block_edges_ms" in conv_ms:
            # closed mesh: signs by propagation. The stage is bounded by instruction issue (fp64 predicates per projected lattice
            # column, bit-parallel flood fill); reported against HBM with the bytes it has to move as an honest lower bound.
            sp_ms = sum(conv_ms.get(k, 0.0) for k in ("sign_block_edges_ms", "sign_components_ms", "sign_brute_ms", "sign_broadcast_ms"))
            nb = work.get("n_bricks", 0.0)
            by = 36.0 * n_tris * 2 + nb * (192 * 2 + 64 + 800) + 8.0 * work.get("n_active", 0.0)
            rl.append({"kernel": "k_block_edges + k_sp_bricks + k_sp_faces + k_sp_flatten + k_sp_stream + k_sp_broadcast (sign propagation)", "bound": "hbm", "achieved": by / (sp_ms * 1e-3) / 1e9,
                       "peak": hbm_peak, "unit": "GB/s", "ms": sp_ms, "traffic": None, "peak_source": hbm_src,
                       "algorithmic": "72 B x n_tris (two passes over the triangles) + 1248 B x n_bricks (edge masks, component tables) + 8 B x n_active (sign write-back); %d representatives evaluated" % int(work.get("n_sign_seeds", 0))})
        if "mc_emit_ms" in mc_ms:
            nb = work.get("n_bricks", 0.0)
            if "mc_count_ms" in mc_ms:
                rl.append({"kernel": "k_mc_count (stage + classify + MC33 tiling + exact triangle count per brick)", "bound": "hbm", "achieved": nb * (2112 + 868) / (mc_ms["mc_count_ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "ms": mc_ms["mc_count_ms"], "algorithmic": "n_bricks x 2980 B (values + masks + halo gathers)", "traffic": None, "peak_source": hbm_src})
            by = nb * (2112 + 868) + 36.0 * (n_verts_local / 3.0) * (world if (world > 1 and args.gather == "fused") else 1)
            rl.append({"kernel": "k_mc_emit (stage + classify + MC33 + vertices, staged coalesced stores%s)" % (" into every rank's buffer" if (world > 1 and args.gather == "fused") else ""), "bound": "hbm", "achieved": by / (mc_ms["mc_emit_ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "ms": mc_ms["mc_emit_ms"], "algorithmic": "n_bricks x 2980 B + 36 B x n_out_tris" + (" x world" if (world > 1 and args.gather == "fused") else ""), "traffic": None, "peak_source": hbm_src})
        if world > 1 and args.gather == "fused" and "mc_emit_ms" in mc_ms:
            # the fused compute + exchange kernel: bytes that must cross NVLink / the measured peer-copy rate of this pool (B200_PROFILING.md)
            by = 36.0 * (n_verts_local / 3.0) * (world - 1)
            rl.append({"kernel": "k_mc_emit as the output exchange (every triangle stored into the %d other ranks' buffers while it is computed)" % (world - 1), "bound": "nvlink",
                       "achieved": by / (mc_ms["mc_emit_ms"] * 1e-3) / 1e9, "peak": 770.0, "unit": "GB/s", "ms": mc_ms["mc_emit_ms"], "traffic": None, "peak_source": "measured peer copy, per direction per GPU (B200_PROFILING.md)",
                       "algorithmic": "36 B x n_out_tris of this rank x (world - 1) outbound"})
        # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum): read from the committed ncu capture of this exact
        # workload (profiles/r2_ncu_traffic_cfg5.json, written by tools/ncu_traffic.py from an `ncu --set full` report)
        if args.config == 5 and args.scale == 1.0 and world == 1:
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic_cfg5.json")))["kernels"]
                for r in rl:
                    k = tr.get(r["kernel"].split(" ")[0])
                    r["traffic"] = k["dram_bytes_per_launch"] if k else None
                    if k:
                        r["traffic_source"] = "profiles/r2_ncu_traffic_cfg5.json"
            except Exception:
                pass
        for r in rl:
            r["frac"] = r["achieved"] / r["peak"]
            if r["bound"] == "fp32":
                r["note"] = "FP32 CUDA-core kernel (no stage of this path is a dense contraction or a stream): peak = SMs x 128 lanes x 2 x max SM clock; achieved counts algorithmic FLOPs only"
        stage_all = dict(conv_ms)
        stage_all.update(mc_ms)
        dominant = max(rl, key=lambda r: r["ms"]) if rl else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config %d: %s" % (args.config, workload_desc(args)), "mesh": desc, "n_triangles": int(n_tris), "voxel_size": vs,
                       "band_width": 0, "l2": "inputs larger than L2 (triangles %.0f MB, bricks %.0f MB)" % (tris.nbytes / 1e6, work.get("n_bricks", 0) * 2112 / 1e6),
                       "parallelism": "brick slabs x%d, mesh replicated" % world, "output_exchange": (args.gather if world > 1 else None), "sign_propagation": bool(work.get("sign_propagation", 0.0))},
            "remesh_ms": ms_per_step, "tris_per_s": (n_verts / 3.0) / (ms_per_step * 1e-3), "n_active_voxels": n_active, "n_out_triangles": n_verts / 3.0,
            "step_ms": step_ms, "step_phase_ms": phase_ms, "stage_ms": stage_all, "stage_ms_per_rank": per_rank, "work": work,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(tris.nbytes), "d2h_bytes_per_step": int(n_verts * 12),
                    "what": ("one bs_voxel_remesh_into call (VoxelRemesher::remesh): pinned host triangles in, all output vertices in pinned host memory; %d slabs, the read-back of one slab overlaps the kernels of the next" % int((e2e_stats or {}).get("remesh_slabs", 0))) if e2e_stats is not None else
                            "pinned host triangles in (N > 1: uploaded by rank 0, broadcast over NVLink), all output vertices back in host memory (N > 1: every rank copies its slice into one shared page-locked buffer)",
                    "stage_ms_sum_over_slabs": {k: v for k, v in (e2e_stats or {}).items() if k.endswith("_ms") and k != "total_ms"} or None},
            "e2e_indexed": None, "gpu_launches": None, "clocks": clocks,
            "roofline": dominant, "rooflines": rl,
            "verified": verify_out["verified"] if verify_out else None, "verify": verify_out,
        }
        if e2e_indexed:
            e2e_indexed["value"] = n_active / (e2e_indexed["ms_per_step"] * 1e-3)
            e2e_indexed["unit"] = UNIT
            out["e2e_indexed"] = e2e_indexed
        out["gpu_launches"] = int(launches) * world  # counted by the library at its launch sites (bs_kernel_launch_count)
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            cpu_scale = args.cpu_scale
            if cpu_scale is None:  # one pass of ~20 s on this host, sized from two measured calibration passes
                cpu_scale, _, _ = cpu_pick_scale(args.config, cores, 1, 25.0)
            ctris, cvs, cdesc = workload(args.config, cpu_scale)
            nv_c, nt_c, dt_c, st_c = cpu_port(ctris, cvs, cores)
            out["cpu_baseline"] = {"value": nv_c / dt_c, "unit": UNIT, "cores": cores, "threads": cores, "kind": "port",
                                   "stage_seconds": {k: round(st_c[k], 3) for k in ("t_subdivide", "t_tree", "t_udf", "t_sign", "t_mc")},
                                   "sample": "config %d at scale %g: %s (%d triangles, %d active voxels), one convert + MC pass in %.1f s" % (args.config, cpu_scale, cdesc, ctris.shape[0], nv_c, dt_c)}
        print(json.dumps(out))
    if world > 1:
        if peer["ptrs"] is not None:  # unmap the peers' buffers before anyone frees its own
            torch.cuda.synchronize()
            for r, q in enumerate(peer["ptrs"]):
                if r != rank:
                    L.bs_ipc_close(ctx._h, C.c_void_p(q))
            dist.barrier()
            L.bs_ipc_free(ctx._h, C.c_void_p(peer["own"]))
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
