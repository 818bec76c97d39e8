#!/usr/bin/env python
"""bench.py -- N1curl S+T assembly throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                      CPU arm: C restatement of the reference algorithm

A step = one pass of the hot path: values of S and T(=M) with PEC row elimination for every row this rank
owns, from device-resident mesh, materials, dof map and CSR pattern (numbering + pattern are one-time set-up,
reported separately in "setup_ms").  Workloads (weak scaling towards BASELINE config 5):
  N=1  UnitCube(119) order 2  (10.1 M tets: the largest BASELINE config whose S+T fit one B200)
  N=2  UnitCube(128), N=4 UnitCube(161), N=8 UnitCube(203) = config 5 (50.2 M tets), rows partitioned by
  DOF-entity ownership, ghost cells duplicated: no data-path collective.
Outputs ONE JSON line on rank 0 (see the task contract): value / e2e / roofline / cpu_baseline / clocks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'sucem-fem_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "N1curl order-2 S+T elements assembled/s (fp64)"
UNIT = "tets/s"
WORKLOAD_N = {1: 119, 2: 128, 4: 161, 8: 203}


def algorithmic_bytes(n_cells, n_vertices, nnz, n_mat=2):
    """SURVEY 8(d) / BASELINE.md section 4: values out + connectivity + coordinates + (eps_r, mu_r)."""
    return 8 * nnz * n_mat + 16 * n_cells + 24 * n_vertices + 16 * n_cells


def peak_hbm():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons sampled DURING the timed region (NVML, ~1 ms period; the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = int(vis.split(',')[self.index]) if vis and vis.split(',')[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksThrottleReasons(h),
                                  nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
                time.sleep(0.001)
        except Exception as e:  # noqa: BLE001
            self.error = repr(e)

    def summary(self):
        import pynvml as nv
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = [('hw_slowdown', nv.nvmlClocksThrottleReasonHwSlowdown),
                 ('hw_thermal_slowdown', nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ('sw_thermal_slowdown', nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                 ('sw_power_cap', nv.nvmlClocksThrottleReasonSwPowerCap)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, b in names if bits & b], "samples": len(self.rows),
                "power_w_max": max((r[2] for r in self.rows), default=None)}


def cpu_port_run(n, threads, steps, warmup=0):
    """C restatement of the reference algorithm on the host cores: returns (tets/s, seconds per step, cells)."""
    import numpy as np
    from oracle import cport, meshes
    co, ce = meshes.unit_cube(n)
    rng = np.random.default_rng(1234)
    eps, mu = rng.uniform(1, 16, len(ce)), rng.uniform(1, 2, len(ce))
    A = cport.CPUAssembler(co, ce, 2)
    S, M = np.zeros(len(A.cols)), np.zeros(len(A.cols))
    ts = []
    for it in range(warmup + steps):
        S[:] = 0
        M[:] = 0
        t0 = time.perf_counter()
        A.assemble(eps, mu, threads=threads, S=S, M=M)
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    return len(ce) / dt, dt, len(ce), A


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host cores (oracle port; DOLFIN is not installable)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import cport  # noqa: F401
    threads = os.cpu_count() or 1
    n = 32
    v, dt, cells, A = cpu_port_run(n, threads, max(args.steps, 1), min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "UnitCube(%d) order-2 S+T PEC, bounded CPU sample of the GPU arm's workload" % n},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": A.max_threads(), "kind": "port",
                             "sample": "UnitCube(%d) order 2, %d tets/step, OpenMP atomics over cells; C restatement of "
                                       "the reference algorithm (DOLFIN 1.0 is not installable here)" % (n, cells)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--n', type=int, default=0, help='override the cube size (debug; not a bench line)')
    ap.add_argument('--kernel', type=int, default=0, help='0 auto, 1 generic, 2 canonical')
    ap.add_argument('--order', type=int, default=2, help='basis order (debug; the bench line is order 2)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world == 1:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', '29517'] + sys.argv
        os.execv(sys.executable, cmd)
    import numpy as np
    import torch
    import torch.distributed as dist
    from sucemfem import device, _native as N
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    n = args.n or WORKLOAD_N.get(world, int(round((world * 203 ** 3 / 8.0) ** (1 / 3.0))))

    # ---- one-time set-up (not timed as part of a step): mesh, numbering, this rank's row block ----------
    t_setup = time.perf_counter()
    mesh = device.DeviceMesh.unit(n, n, n)
    space = device.DeviceSpace(mesh, args.order)
    bounds = space.partition(world)
    pat = device.DevicePattern(space, bounds[rank], bounds[rank + 1])
    markers = torch.zeros(space.n_facets, dtype=torch.int32, device='cuda')
    space.mark_exterior_facets(markers, 999)
    flags = space.pec_row_flags(markers, 999)
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t_setup) * 1e3
    C_total = mesh.num_cells
    g = torch.Generator(device='cuda')
    g.manual_seed(1234)
    eps = 1.0 + 15.0 * torch.rand(C_total, dtype=torch.float64, device='cuda', generator=g)
    mu = 1.0 + torch.rand(C_total, dtype=torch.float64, device='cuda', generator=g)
    S, T = pat.new_values(), pat.new_values()
    kernel = args.kernel

    def step():
        pat.assemble_ST(eps, mu, flags, S, T, kernel=kernel)

    for _ in range(W):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ms_local = ms
    if sampler:
        sampler.stop_flag = True
    per_rank_ms = [ms_local / K]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x.item()) / K for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / K
    value = C_total / (ms_step * 1e-3)

    # ---- roofline: this rank's step (its launches back to back on the launching stream, CUDA events) ------
    peak, peak_src = peak_hbm()
    share = pat.n_rows / float(space.ndof)
    b_alg = algorithmic_bytes(C_total * share, mesh.num_vertices * share, pat.nnz)
    achieved = b_alg / (ms_local / K * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes per step from the committed ncu capture of this workload (profiles/), if it is this workload
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            tj = json.load(f)
        if tj.get('n') == n and world == 1:
            traffic = tj['dram_bytes_per_step']
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg,
                "kernel_ms": ms_local / K,
                "note": "one step = the launches of one sfem_assemble_ST call (k_canon_edges + k_canon_faces), timed "
                        "back to back with CUDA events on the launching stream; achieved = algorithmic bytes of the "
                        "step / step time; traffic = ncu DRAM bytes of the same two launches"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "UnitCube(%d) N1curl order-2 S+T, PEC on all exterior faces, per-cell eps_r~U[1,16) "
                                   "mu_r~U[1,2)" % n, "tets": C_total, "ndof": space.ndof, "nnz_per_matrix_this_rank": pat.nnz,
                       "partition": "rows by DOF-entity ownership, %d block(s), ghost cells duplicated" % world,
                       "l2": "outputs per step (%.1f GB) exceed L2; no flush needed" % (16e-9 * pat.nnz),
                       "kernel": {0: "auto", 1: "generic", 2: "canonical"}[kernel]},
            "setup_ms": setup_ms, "roofline": roofline, "gpu_launches": None,
            "per_rank_ms_per_step": [round(x, 3) for x in per_rank_ms]}

    # ---- e2e through the public API with HOST buffers (rank-local; N=1 only measures the full pipeline) ---
    if not args.no_e2e and world == 1:
        line["e2e"] = e2e_run(n, min(K, 3))
    # ---- CPU baseline (rank 0, N=1): bounded sample, 1 thread = what sucemfem + DOLFIN 1.0 run ------------
    if rank == 0 and world == 1 and not args.no_cpu:
        v1, dt1, cells, A = cpu_port_run(24, 1, 2)
        vall, dtall, _, _ = cpu_port_run(24, A.max_threads(), 2, 1)
        line["cpu_baseline"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "UnitCube(24) order 2 (%d tets) x2 runs, serial cell loop + binary-search "
                                          "insertion; same code with %d OpenMP threads: %.0f tets/s" % (cells, A.max_threads(), vall)}
    if rank == 0:
        line["clocks"] = sampler.summary()
        line["gpu_launches"] = launches_per_step(kernel, pat) * K
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def launches_per_step(kernel, pat):
    from sucemfem import _native as N
    return int(N.lib.sfem_launches_per_assemble_ST(pat.handle, kernel))


def e2e_run(n, steps):
    """The user-facing call with host buffers: numpy mesh + materials in, scipy-ready CSR (rowptr, cols, S, T)
    out, everything between (H2D, numbering, pattern, values, PEC, D2H) inside the timed region."""
    import numpy as np
    import torch
    from sucemfem import device
    gen = device.DeviceMesh.unit(n, n, n)          # synthesise the HOST input arrays (outside the timed region)
    co, ce = gen.coords.cpu().numpy(), gen.cells.cpu().numpy()
    del gen
    rng = np.random.default_rng(1234)
    eps = rng.uniform(1, 16, len(ce))
    mu = rng.uniform(1, 2, len(ce))
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_co, h_ce, h_eps, h_mu = pin(co), pin(ce), pin(eps), pin(mu)
    out = {}
    times = []
    for it in range(steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d_co, d_ce = h_co.cuda(non_blocking=True), h_ce.cuda(non_blocking=True)
        d_eps, d_mu = h_eps.cuda(non_blocking=True), h_mu.cuda(non_blocking=True)
        mesh = device.DeviceMesh(d_co, d_ce)
        space = device.DeviceSpace(mesh, 2)
        pat = space.pattern()
        markers = torch.zeros(space.n_facets, dtype=torch.int32, device='cuda')
        space.mark_exterior_facets(markers, 999)
        flags = space.pec_row_flags(markers, 999)
        S, T = pat.assemble_ST(d_eps, d_mu, flags)
        rp, cl = pat.csr()
        if not out:
            out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in
                   (('S', S), ('T', T), ('rowptr', rp), ('cols', cl))}
            torch.cuda.synchronize()
            t0 = time.perf_counter()  # first iteration (allocating pinned buffers) is the warm-up
        out['S'].copy_(S, non_blocking=True)
        out['T'].copy_(T, non_blocking=True)
        out['rowptr'].copy_(rp, non_blocking=True)
        out['cols'].copy_(cl, non_blocking=True)
        torch.cuda.synchronize()
        if it > 0:
            times.append(time.perf_counter() - t0)
        del S, T, rp, cl, pat, space, mesh
    dt = sorted(times)[len(times) // 2]      # median step
    h2d = sum(x.numel() * x.element_size() for x in (h_co, h_ce, h_eps, h_mu))
    d2h = sum(x.numel() * x.element_size() for x in out.values())
    return {"value": len(ce) / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": dt * 1e3, "steps_ms": [round(t * 1e3, 1) for t in times],
            "includes": "H2D mesh+materials, numbering, pattern, S+T values, PEC, D2H CSR"}


if __name__ == '__main__':
    main()
