#!/usr/bin/env python
"""Benchmark of the PyMOC time-stepping hot path (BASELINE.json metric: member-timesteps/sec, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl reference]

One bench "step" = one fused-kernel pass advancing every member of the ensemble by
``--nt`` model time steps (default: the workload's N from SURVEY.md section 8d).  Under
torchrun every rank holds ``--members`` members (weak scaling) and steps them with no
collective; a final NCCL all-gather of one diagnostic per member runs outside the timed
region.  Prints ONE JSON line (rank 0).

Keys beyond the base contract:
  roofline     FP64-pipe roofline of the fused kernel: algorithmic flops per member-step
               (restated in DESIGN.md) x member-steps/s, against the DFMA
               peak measured live by pmoc_fp64_peak on the same GPU.
  cpu_baseline the oracle's reference-faithful loop (NumPy/SciPy, oracle/pymoc_oracle.py) timed
               on the host cores of this box on a bounded sample of the same workload.
  e2e          the same metric through pmoc_model_run_host with pinned HOST buffers: H2D of
               state + parameters, fused kernel, D2H of state + diagnostics, every step.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pymoc_b200 import configs  # noqa: E402

WORKLOADS = {
    # name: (builder(M) -> spec, default members per GPU, default model steps per bench step)
    'C1': (lambda M: configs.c1_timestepping(M), 16384, 3000),
    'C2': (lambda M: configs.c2_column_so(M), 65536, 7200),
    'twocol': (lambda M: configs.twocol(M), 32768, 2400),
    'C3': (lambda M: configs.c3_twocol_so(M), 32768, 2400),
    # the literal script: F2010 smoother of Psi_GM (c = 0.1, bvp_with_Ek), parity at the stated 1e-5
    'C3_bvp': (lambda M: configs.c3_twocol_so(M, c=0.1), 32768, 2400),
    'twobasin': (lambda M: configs.twobasin(M), 32768, 2400),
    'C4': (lambda M: configs.c4_jansen_nadeau(M), 32768, 2400),
    'C5': (lambda M: configs.c5_single_global_basin(M), 32768, 2400),
    # BASELINE configs[4]: nz=4096 (block-per-member kernels), stable dt = 0.01 d, K = 72 000
    'C5_4096': (lambda M: configs.c5_single_global_basin(M, nz=4096, dt_days=0.01, kapfac_max=1.), 16384, 720),
}


def algorithmic_flops(spec):
  """FP64 flops per member-timestep, SURVEY.md section 8d (add/sub/mul/div/cmp = 1, FMA = 2)."""
  n, K, B = spec.nz, spec.K, spec.nb
  per_step = 11 * n + (2 * n if spec.basin.do_conv else 0)
  ncol = 1
  if spec.north is not None:
    per_step += 11 * n + (2 * n if spec.north.do_conv else 0)
    ncol = 2
  nclos = 1
  if getattr(spec, 'pac', None) is not None:  # third column, second thermal wind + remap, second Psi_SO
    per_step += 11 * n + (2 * n if spec.pac.do_conv else 0)
    ncol, nclos = 3, 2
  refresh = 3 * n * ncol
  if spec.tw is not None:
    refresh += nclos * 16 * n
    if spec.iso:
      refresh += nclos * (6 * B * (n - 1) + 8 * n + 2 * B + 2 * n * (math.ceil(math.log2(B)) + 5))
  if spec.so is not None:
    refresh += nclos * n * (math.ceil(math.log2(spec.so.y.size)) + 20)
    if spec.so.c is not None:
      refresh += 120 * n  # F2010 smoother (approximate)
  if spec.ml is not None:
    m = spec.ml.y.size
    per_step += m * (math.ceil(math.log2(n)) + 32) + n
  return per_step + refresh / K


# DRAM bytes (read + write) of one fused-kernel launch, from the committed `ncu --set full` captures
# (profiles/r1final_full_summary.txt).  State and parameters are read once and state + diagnostics written once
# per launch whatever the number of steps, so the 720/240-step captures stand for the bench's launches.
NCU_TRAFFIC = {  # workload -> (members in the capture, bytes)
    'C2': (65536, 421.1e6 + 362.5e6),
    'C3': (32768, 232.2e6 + 386.6e6),
    'C4': (32768, 257.4e6 + 478.5e6),
}


# ---------------------------------------------------------------------------- CPU reference
def _cpu_task(args):
  workload, M, member, nsteps = args
  os.environ['OMP_NUM_THREADS'] = '1'
  import warnings
  warnings.filterwarnings('ignore')
  from oracle import pymoc_oracle as O
  spec = WORKLOADS[workload][0](M)
  case = spec.member_case(member)
  t0 = time.perf_counter()
  O.run_coupled(case, nsteps, O.REFERENCE)
  return time.perf_counter() - t0


def cpu_reference(workload, M, nsteps=0, tasks_per_core=2, cores=None, target_s=15.0):
  """Reference-faithful CPU loop (the oracle in its default modes: solve_bvp, brentq, dense inv
  exactly where the reference calls them) on all host cores; returns member-steps/s.  With
  ``nsteps == 0`` the steps per task are sized from a short pilot so that the sample takes about
  ``target_s`` seconds of wall time (a bounded sample of the same workload)."""
  import multiprocessing as mp
  cores = cores or os.cpu_count() or 1
  ntask = cores * tasks_per_core
  members = [int(i * (M - 1) / max(ntask - 1, 1)) for i in range(ntask)]
  K = WORKLOADS[workload][0](min(M, 64)).K
  ctx = mp.get_context('spawn')
  with ctx.Pool(cores) as pool:
    pilot_steps = 2 * K + 1
    t_pilot = max(pool.map(_cpu_task, [(workload, M, m, pilot_steps) for m in members[:cores]]))  # import + warm-up
    if nsteps <= 0:
      per_step = t_pilot / pilot_steps
      nsteps = int(max(pilot_steps, min(200000, target_s / (per_step * tasks_per_core))))
      nsteps = max(K, (nsteps // K) * K) + 1
    jobs = [(workload, M, m, nsteps) for m in members]
    t0 = time.perf_counter()
    pool.map(_cpu_task, jobs, chunksize=1)
    wall = time.perf_counter() - t0
  return ntask * nsteps / wall, cores, '%d members x %d steps of %s (members spread over the lattice)' % (
      ntask, nsteps, workload), wall


# ------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
  Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.rows, self.proc = index, [], None

  def run(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
      for line in self.proc.stdout:
        self.rows.append([c.strip() for c in line.split(',')])
    except Exception:
      pass

  def stop(self):
    if self.proc is not None:
      self.proc.terminate()
    self.join(timeout=2)
    sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
    names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
    reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == 'active'})
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': reasons, 'samples': len(sm)}


# ------------------------------------------------------------------------------ GPU arm
class PinnedHostBackend:
  """Host-side twin of CudaBackend for the e2e leg: pinned host buffers whose addresses go to
  pmoc_model_run_host (which does the H2D / D2H itself)."""

  def __init__(self, lib):
    import torch
    self.torch, self.lib, self.bytes_in, self.bytes_out = torch, lib, 0, 0

  def upload(self, arr):
    a = np.ascontiguousarray(arr)
    t = self.torch.empty(a.shape, dtype=self.torch.from_numpy(a[:0]).dtype, pin_memory=True)
    t.numpy()[...] = a
    self.bytes_in += a.nbytes
    return t

  def zeros(self, shape, dtype=np.float64):
    tdt = {np.float64: self.torch.float64, np.int32: self.torch.int32, np.uint32: self.torch.int32}[dtype]
    t = self.torch.zeros(shape, dtype=tdt, pin_memory=True)
    self.bytes_out += t.numel() * t.element_size()
    return t

  @staticmethod
  def ptr(buf):
    return None if buf is None else buf.data_ptr()

  def download(self, buf):
    return buf.numpy().copy()

  def assign(self, buf, arr):
    buf.numpy()[...] = arr

  def stream(self):
    return None

  def sync(self):
    pass


def gpu_arm(args):
  import torch
  import torch.distributed as dist
  from pymoc_b200 import _abi, _lib
  from pymoc_b200.ensemble import Ensemble
  from pymoc_b200.parallel import gather_members, shard_range

  rank = int(os.environ.get('RANK', 0))
  world = int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(local)
  if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
    os.environ['NCCL_DEBUG'] = 'WARN'  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  lib = _lib.lib()

  build, m_default, nt_default = WORKLOADS[args.workload]
  m_local = args.members or m_default
  nt = args.nt or nt_default
  M = m_local * world
  lo, hi = shard_range(M, rank, world)
  with configs.members(lo, hi):  # only this rank's block of the M-member lattice is ever built
    spec = build(M)
  ens = Ensemble(spec)

  peak = ctypes.c_double()
  mhz = ctypes.c_double()
  _lib.check(lib.pmoc_fp64_peak(ctypes.byref(peak), ctypes.byref(mhz), None))

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(args.warmup):
    ens.run(nt, sync=False)
  barrier()
  sampler = ClockSampler(local)
  sampler.start()
  time.sleep(0.3)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  for _ in range(args.steps):
    ens.run(nt, sync=False)
  e1.record()
  barrier()
  ms = e0.elapsed_time(e1)
  clocks = sampler.stop()
  t = torch.tensor([ms], dtype=torch.float64, device='cuda')
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms = float(t.item())
  ms_per_step = ms / args.steps
  value = M * nt / (ms_per_step * 1e-3)

  # final diagnostic gather (the only collective of the path), untimed
  diag_key = 'Psi_so' if spec.so is not None else 'Psi_tw'
  per_member = ens.buffer(diag_key).abs().amax(dim=1)
  gathered = gather_members(per_member, M)
  status = ens.diagnostics()['status']
  # members whose parameters drive the (explicit, CFL-limited) model itself unstable are flagged by the
  # kernel; they cost the same instructions as the others.  The lattices are chosen so that there are none.
  nan_members = int((status & 1).astype(bool).sum())
  assert gathered.shape[0] == M and nan_members <= 1e-3 * ens.M, '%d of %d members non-finite' % (nan_members, ens.M)

  # e2e through the host-buffer C-ABI call
  e2e = None
  if args.e2e_steps > 0:
    hb = PinnedHostBackend(lib)
    hens = Ensemble(spec, backend=hb)
    _lib.check(lib.pmoc_model_run_host(ctypes.byref(hens.model), 0, nt))  # warm-up (allocator pools)
    barrier()
    t0 = time.perf_counter()
    it = nt
    for _ in range(args.e2e_steps):
      _lib.check(lib.pmoc_model_run_host(ctypes.byref(hens.model), it, nt))
      it += nt
    barrier()
    dt_e2e = time.perf_counter() - t0
    tt = torch.tensor([dt_e2e], dtype=torch.float64, device='cuda')
    if world > 1:
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    h2d, d2h = ctypes.c_uint64(), ctypes.c_uint64()
    lib.pmoc_host_last_bytes(ctypes.byref(h2d), ctypes.byref(d2h))  # what the last call actually copied
    e2e = {'value': M * nt * args.e2e_steps / float(tt.item()), 'unit': 'member-timesteps/s',
           'h2d_bytes_per_step': int(h2d.value), 'd2h_bytes_per_step': int(d2h.value),
           'steps': args.e2e_steps, 'api': 'pmoc_model_run_host (pinned host buffers)'}
    hfin = hens.state()['b_basin']
    assert (~np.isfinite(hfin).all(axis=1)).sum() <= 1e-3 * hens.M, 'non-finite members in the host-buffer run'

  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu:
    v, cores, sample, wall = cpu_reference(args.workload, min(M, 65536), args.cpu_steps)
    cpu = {'value': v, 'unit': 'member-timesteps/s', 'cores': cores, 'kind': 'port', 'sample': sample,
           'wall_s': round(wall, 2)}

  if rank == 0:
    flops = algorithmic_flops(spec)
    achieved = value / world * flops / 1e12
    line = {
        'metric': 'member-timesteps/sec, fp64', 'value': value, 'unit': 'member-timesteps/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s: %s' % (args.workload, spec.name), 'members_per_gpu': m_local, 'members': M,
                   'nz': spec.nz, 'ny': spec.ny, 'K': spec.K, 'dt_days': spec.dt / 86400.,
                   'model_steps_per_bench_step': nt, 'nan_members_rank0': nan_members, 'sweep_rank0': {k: [float(v.min()), float(v.max())] for k, v in spec.sweep.items()},
                   'l2': 'inputs (state + per-member parameters, %.0f MB per GPU) larger than the 126 MB L2'
                         % (ens.M * spec.nz * 8 * 3 / 1e6),
                   'parallelism': 'ensemble members sharded over %d GPU(s), no collective in the loop' % world},
        'roofline': {'bound': 'fp64',
                     'bound_note': 'FP64 pipe (DFMA): the state stays on chip across the fused steps, BASELINE.json asks '
                                   'for "% of FP64/HBM roofline"; hbm_check shows the HBM side',
                     'achieved': achieved, 'peak': peak.value, 'unit': 'TFLOP/s',
                     'frac': achieved / peak.value,
                     'traffic': (NCU_TRAFFIC[args.workload][1] * m_local / NCU_TRAFFIC[args.workload][0]
                                 if args.workload in NCU_TRAFFIC else None),
                     'traffic_source': 'bytes per launch (dram read + write), profiles/r1final_full_summary.txt',
                     'hbm_check': (None if args.workload not in NCU_TRAFFIC else {
                         'achieved_GBs': NCU_TRAFFIC[args.workload][1] * m_local / NCU_TRAFFIC[args.workload][0]
                                         / (ms_per_step * 1e-3) / 1e9,
                         'peak_GBs': 6540.5, 'note': 'MEASURED_PEAKS.json hbm_gbs; state stays on chip, HBM is idle'}),
                     'flops_per_member_step': flops, 'peak_source': 'pmoc_fp64_peak measured live (DFMA stream)',
                     'kernel_ms': ms_per_step},
        'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': args.steps, 'clocks': clocks,
    }
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def reference_arm(args):
  rank = int(os.environ.get('RANK', 0))
  world = int(os.environ.get('WORLD_SIZE', 1))
  if rank != 0:
    return
  build, m_default, nt_default = WORKLOADS[args.workload]
  M = (args.members or m_default) * world
  spec = build(min(M, 65536))
  vals = []
  for _ in range(args.warmup):
    cpu_reference(args.workload, min(M, 65536), args.cpu_steps, tasks_per_core=1, target_s=4.0)
  t0 = time.perf_counter()
  for _ in range(args.steps):
    v, cores, sample, wall = cpu_reference(args.workload, min(M, 65536), args.cpu_steps, tasks_per_core=1, target_s=12.0)
    vals.append(v)
  total = time.perf_counter() - t0
  value = float(np.mean(vals))
  line = {
      'impl': 'reference', 'metric': 'member-timesteps/sec, fp64', 'value': value, 'unit': 'member-timesteps/s',
      'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total / args.steps * 1e3,
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': '%s: %s' % (args.workload, spec.name), 'nz': spec.nz, 'ny': spec.ny, 'K': spec.K},
      'cpu_baseline': {'value': value, 'unit': 'member-timesteps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
      'e2e': {'value': value, 'unit': 'member-timesteps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--workload', default='C2', choices=sorted(WORKLOADS))
  ap.add_argument('--members', type=int, default=0, help='members per GPU (default: workload size)')
  ap.add_argument('--nt', type=int, default=0, help='model time steps per bench step')
  ap.add_argument('--e2e-steps', type=int, default=3)
  ap.add_argument('--cpu-steps', type=int, default=0, help='model steps per CPU-baseline task (0: sized for ~15 s)')
  ap.add_argument('--no-cpu', action='store_true')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
  if args.impl == 'reference':
    reference_arm(args)
  else:
    gpu_arm(args)


if __name__ == '__main__':
  main()
