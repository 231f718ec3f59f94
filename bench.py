#!/usr/bin/env python
"""Benchmark of the ciMRGP VI hot path (BASELINE.json: VI iterations/sec at N = 1e6, 10 resolutions).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one full variational sweep (`_fit()`, MRGP.py:571-652) over all 10 layers of the config-4
workload (N = 1e6 samples, dx = 1, dy = 2, M = 30, 1023 regions, ci mode, fp64, static basis intervals).
Prints ONE JSON line (see the keys below).  Timing: CUDA events on the engine's stream around every step,
L2 flushed (256 MB write) between steps, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np


def _finite(o):
    """JSON has no NaN / inf: map them to null."""
    if isinstance(o, dict):
        return {k: _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    if isinstance(o, float) and not np.isfinite(o):
        return None
    return o


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

N_SAMPLES = 1000000
N_LAYERS = 10
N_BASIS = 30
DY = 2
WORKLOAD = 'config4: synthetic 1-D nonstationary signal (script-1 f, seed 10), N=1e6, dx=1, dy=2, M=30, ' \
           '10 resolutions (1023 regions), ciMRGP, fp64, static basis intervals'
# algorithmic HBM bytes per sample-layer: phase A 8(dx+2dy) = 40, phase B 8(dx+2dy+1) + 8(dy+1) = 72
BYTES_SWEEP_PER_SAMPLE_LAYER = 112


def peak_hbm():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a polling thread (one sample
    every ~2 ms; the timed region of this benchmark lasts tens of milliseconds, too short for `nvidia-smi -lms`),
    with the nvidia-smi one-shot query of B200_PROFILING.md as the fallback."""
    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        self.index, self.sm, self.mask, self.max_sm, self.stop_flag, self.thread, self.nvml = index, [], 0, None, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)))
                get = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.mask |= int(get(self.dev))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.max_sm,
                    'reasons': [n for n, bit in self.REASONS if self.mask & bit], 'samples': len(self.sm), 'source': 'nvml'}
        try:
            q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
                'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
            out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(',')
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            return {'sm_mhz': float(out[0]), 'sm_max_mhz': float(out[1]),
                    'reasons': [n for k, n in enumerate(names) if out[2 + k].strip() == 'Active'], 'samples': 1,
                    'source': 'nvidia-smi after the timed region'}
        except Exception:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml and nvidia-smi unavailable'], 'samples': 0}


def make_model(n, device, n_ctas=0, distributed=False):
    import workloads
    from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel
    from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
    x, y = workloads.workload1(n)
    m = MultiResolutionGaussianProcess([x, y], N_BASIS, IndexSetUniform(n, N_LAYERS - 1, 2), LaplacianEigenpairs(),
                                       MaternKernel(nu=1, l=1, sf=1), forced_independence=False, device=device,
                                       n_ctas=n_ctas, distributed=distributed)
    return m


def oracle_sample(n_sample, sweeps):
    """CPU oracle (NumPy port of the reference) on a bounded sample of the workload: the same signal,
    layers and basis at n_sample samples.  Returns seconds per sweep and the fixed (N-independent) part
    spent in the permutation-weight solver."""
    import workloads
    from oracle import mrgp_oracle as O
    x, y = workloads.workload1(n_sample)
    m = O.OracleMRGP(x, y, N_BASIS, O.uniform_offsets(n_sample, N_LAYERS - 1, 2), mode='ci')
    times = []
    for _ in range(sweeps):
        t0 = time.perf_counter()
        m.sweep()
        times.append(time.perf_counter() - t0)
    return times, m.t_omega / sweeps


def scaled_rate(t_step, t_fixed, n_sample):
    """it/s on the full N from a step on n_sample samples: streaming part scales linearly in N, the
    permutation-weight solve does not depend on N."""
    return 1.0 / (t_fixed + (t_step - t_fixed) * (float(N_SAMPLES) / n_sample))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (its NumPy port, oracle/; the
    reference itself is Python and does not exist on the GPU box) on the host cores."""
    if rank != 0:
        return
    budget = 150.0
    per_step = budget / max(1, args.steps + args.warmup)
    n_sample = int(min(N_SAMPLES, max(20000, (per_step - 1.5) / 41.5e-6)))
    n_sample = (n_sample // 1024) * 1024
    times, t_fixed = oracle_sample(n_sample, args.steps + args.warmup)
    t_step = float(np.mean(times[args.warmup:]))
    value = scaled_rate(t_step, t_fixed, n_sample)
    sample = '%d steps of one oracle sweep on N=%d samples (same 10 layers, M=30); scaled to N=1e6 as ' \
             '1/(t_omega + (t_step - t_omega) * 1e6/N), t_step=%.2fs, t_omega=%.2fs' % (args.steps, n_sample, t_step, t_fixed)
    line = {
        'impl': 'reference', 'metric': 'ciMRGP VI iters/sec at N=1e6, R=10', 'value': value, 'unit': 'it/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 / value,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD},
        'cpu_baseline': {'value': value, 'unit': 'it/s', 'cores': 1, 'kind': 'port', 'sample': sample,
                         'host_cores': os.cpu_count()},
        'e2e': {'value': value, 'unit': 'it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(_finite(line)))


def time_phase(eng, fn, j, torch, reps):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for _ in range(reps):
        ev0.record(eng.stream)
        fn(j)
        ev1.record(eng.stream)
        ev1.synchronize()
        out.append(ev0.elapsed_time(ev1))
    return out


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    # strong scaling: the N = 1e6 problem is split into contiguous sample chunks, one per GPU
    m = make_model(N_SAMPLES, local_rank, args.ctas, distributed=multi)
    eng = m._engine
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    peak, peak_src = peak_hbm()

    def barrier():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if not multi:
            return v
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        eng.sweep(1)
    eng.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing: K sweeps, L2 flushed between steps ------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = eng.launch_count()
    barrier()
    for k in range(args.steps):
        # the L2 flush (256 MB of HBM writes on torch's stream) must neither overlap the sweep before it nor the one
        # after it: it starts when the previous timed sweep is over and ends before the next timed region starts
        torch.cuda.current_stream().wait_stream(eng.stream)
        flush.zero_()
        eng.stream.wait_stream(torch.cuda.current_stream())
        barrier()
        ev[k][0].record(eng.stream)
        eng.sweep(1)
        ev[k][1].record(eng.stream)
    barrier()
    launches = eng.launch_count() - l0
    exchange = getattr(eng, 'exchange', None)
    if multi and exchange == 'nccl':
        launches = 80 * args.steps   # per rank and sweep: 10 x (phase A, sums, mid, omega, phase B, sums, bias/noise) + NCCL
    step_ms = [max_over_ranks(a.elapsed_time(b)) for a, b in ev]
    total_s = sum(step_ms) / 1e3
    value = args.steps / total_s

    # ---- end to end through the public API: pinned host -> device, sweep, ELBO terms back ----------
    e2e = None
    if not args.no_e2e:
        e2e_ms = []
        n_local = eng.N
        for k in range(args.steps):
            flush.zero_()
            eng.stream.wait_stream(torch.cuda.current_stream())
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(eng.stream)
            eng.upload_observations()          # H2D of this rank's rows of y (n,2) from pinned memory
            m.fit(n_iter=1, tol=1e-300, min_iter=1)    # one sweep + the six ELBO terms per layer, read back
            b.record(eng.stream)
            b.synchronize()
            e2e_ms.append(max_over_ranks(a.elapsed_time(b)))
        e2e = {'value': args.steps / (sum(e2e_ms) / 1e3), 'unit': 'it/s',
               'h2d_bytes_per_step': n_local * DY * 8, 'd2h_bytes_per_step': N_LAYERS * 6 * 8,
               'ms_per_step': float(np.mean(e2e_ms)), 'lower_bound_layer0': m.lower_bound_layer[0][-1]}
    clocks = sampler.stop() if rank == 0 else None
    if multi:
        if rank == 0:
            line = {
                'metric': 'ciMRGP VI iters/sec at N=1e6, R=10', 'value': value, 'unit': 'it/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': 1e3 * total_s / args.steps,
        'step_ms': [round(v, 4) for v in step_ms],
                'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'l2': 'flushed between timed steps (256 MB write)',
                           'parallelism': ('samples sharded in %d contiguous chunks; 2 exchanges of <= 245 KB of region '
                                           'statistics per layer, ' % world) +
                                          ('own kernels over NVLink peer memory, whole sweep in one CUDA graph'
                                           if exchange == 'peer' else 'NCCL all-reduce'),
                           'exchange': exchange},
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches,
                'roofline': {'bound': 'hbm', 'achieved': BYTES_SWEEP_PER_SAMPLE_LAYER * N_SAMPLES * N_LAYERS / (total_s / args.steps) / 1e9,
                             'peak': peak * world, 'peak_source': peak_src, 'unit': 'GB/s',
                             'frac': BYTES_SWEEP_PER_SAMPLE_LAYER * N_SAMPLES * N_LAYERS / (total_s / args.steps) / 1e9 / (peak * world),
                             'traffic': None, 'note': 'whole-sweep algorithmic bytes (112 B per sample-layer) over all GPUs'},
                'cpu_baseline': None,
            }
            print(json.dumps(_finite(line)))
            sys.stdout.flush()
        # leave without running destructors: tearing down a process group whose collectives live in a captured CUDA
        # graph can block at interpreter exit
        barrier()
        sys.stdout.flush()
        os._exit(0)

    # ---- per-kernel timing for the roofline ----------------------------------------------------------
    # The captured ci sweep streams the samples on layer 0 only (phase A and phase B over x and y); the layers above
    # take their statistics in closed form and every other kernel is a latency-bound small-matrix step.
    # Phase A of layer 0 is the same kernel in the sweep and behind mrgp_phase_a: it is timed alone with CUDA events
    # on the engine stream, L2 flushed before every launch.
    import ctypes as C
    dom_samples = []
    for _ in range(7):
        flush.zero_()
        eng.stream.wait_stream(torch.cuda.current_stream())
        dom_samples += time_phase(eng, eng.phase_a, 0, torch, 1)
    dom_ms = float(np.median(dom_samples))
    dom_bytes = 24 * N_SAMPLES           # x (8) + y (16) per sample; layer 0 has no latent input
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # where the time of one sweep goes: global-timer stamps written by the kernels inside the captured graph
    eng.lib.mrgp_timeline_enable(eng.handle, 1)
    eng.sweep(2)
    eng.synchronize()
    tags, tms = (C.c_int32 * 128)(), (C.c_float * 256)()
    eng.lib.mrgp_timeline_read(eng.handle, tags, tms, 128)
    eng.lib.mrgp_timeline_enable(eng.handle, 0)
    def stamp(j, k):
        b, e = tms[2 * (4 * j + k)], tms[2 * (4 * j + k) + 1]
        return None if b < 0 else [round(1e3 * b, 1), round(1e3 * e, 1)]
    timeline = {name: [stamp(j, k) for j in range(N_LAYERS)]
                for k, name in enumerate(('phase_a', 'mid', 'phase_b_or_closed_form', 'omega_side_stream'))}
    ends = [v[1] for rows in timeline.values() for v in rows if v]
    sweep_us = max(ends) if ends else None
    stream_us = sum(v[1] - v[0] for v in (timeline['phase_a'][0], timeline['phase_b_or_closed_form'][0]) if v)
    roofline = {'bound': 'hbm', 'kernel': 'k_phase_a<2,30,observed targets,no latent input> (layer 0)', 'achieved': achieved,
                'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': 24041216, 'traffic_source': 'ncu --set full, profiles/r01_ncu_summary.md',
                'bytes_per_launch': dom_bytes, 'ms_per_launch': dom_ms,
                'note': 'FP64-bound kernel (see fp64); the sweep is dominated by latency-bound small-matrix kernels',
                'sweep_timeline_us': timeline, 'sweep_us_in_graph_warm_l2': sweep_us,
                'streaming_share_of_sweep': (stream_us / sweep_us) if sweep_us else None}
    # FP64 pipe: measured DFMA peak next to the FMA count of that kernel (6 per basis function and sample:
    # recurrence 1 and Phi A 2 in the first pass, recurrence 1 and Phi^T r 2 in the second; + ~40 for sincospi)
    sink = torch.zeros(8, dtype=torch.float64, device='cuda')
    ms = C.c_float()
    iters = 200000
    torch.cuda.synchronize()
    eng.lib.mrgp_fp64_probe(None, iters, C.c_void_p(sink.data_ptr()), C.byref(ms))
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp64_peak = sms * 4 * 256 * iters * 8 * 2 / (ms.value * 1e-3) / 1e12
    dom_flops = (6 * N_BASIS + 40) * 2 * N_SAMPLES
    roofline['fp64'] = {'peak_tflops_measured': fp64_peak, 'achieved_tflops': dom_flops / (dom_ms * 1e-3) / 1e12,
                        'frac': dom_flops / (dom_ms * 1e-3) / 1e12 / fp64_peak}

    cpu = None
    if not args.no_cpu_baseline:
        n_sample = 200000
        times, t_fixed = oracle_sample(n_sample, 1)
        cpu_value = scaled_rate(times[0], t_fixed, n_sample)
        cpu = {'value': cpu_value, 'unit': 'it/s', 'cores': 1, 'kind': 'port', 'host_cores': os.cpu_count(),
               'sample': 'one oracle sweep (NumPy port of MRGP._fit) on N=%d of the same signal, 10 layers, M=30: '
                         '%.1fs, of which %.1fs in the N-independent fsolve; scaled to N=1e6 as '
                         '1/(t_omega + (t - t_omega) * 5)' % (n_sample, times[0], t_fixed)}

    line = {
        'metric': 'ciMRGP VI iters/sec at N=1e6, R=10', 'value': value, 'unit': 'it/s', 'n_gpus': 1,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': 1e3 * total_s / args.steps,
        'step_ms': [round(v, 4) for v in step_ms],
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'l2': 'flushed between timed steps (256 MB write)',
                   'n_ctas': eng.lib and args.ctas or 'one persistent CTA per SM'},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu,
        'omega_iters_last_sweep': [int(v) for v in eng.get(-1, 51, (N_LAYERS,))],
        'batched_cholesky': {'count_total': eng.cholesky_count(), 'n': DY,
                             'note': 'dy x dy PD guard inside k_mid2; < 1% of a sweep'},
    }
    print(json.dumps(_finite(line)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--ctas', type=int, default=0)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
