"""Multi-rank path on CPU: world_size-2 gloo, one process per "GPU" (SURVEY.md section 8e).

Each rank builds only its member block (shard_range), steps it through the kernel sources on the
warp emulator with no collective in the loop, and the diagnostics are gathered at the end with the
same gather_members that runs over NCCL on the B200s.  The gathered result must equal a single
process stepping the whole ensemble, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, world, port, M, nsteps, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  for p in (os.path.dirname(HERE), HERE):
    if p not in sys.path:
      sys.path.insert(0, p)
  import torch
  import torch.distributed as dist
  from emu.emu_backend import EmuBackend
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  from pymoc_b200.parallel import gather_members, shard_range
  dist.init_process_group('gloo', rank=rank, world_size=world)
  spec = configs.c3_twocol_so(M, axes=(3, 2, 2, 1) if M == 12 else None)
  lo, hi = shard_range(M, rank, world)
  ens = Ensemble(spec, backend=EmuBackend(), members=(lo, hi))
  ens.run(nsteps)
  b = gather_members(torch.from_numpy(ens.state()['b_basin']), M)
  psi = gather_members(torch.from_numpy(ens.diagnostics()['Psi_iso_b']), M)
  assert b.shape[0] == M and psi.shape[0] == M
  if rank == 0:
    np.savez(os.path.join(out_dir, 'gathered.npz'), b=b.numpy(), psi=psi.numpy())
  dist.barrier()
  dist.destroy_process_group()


def test_shard_range_partitions():
  from pymoc_b200.parallel import shard_range
  for M in (1, 7, 8, 65536, 262144 + 3):
    for world in (1, 2, 3, 8):
      blocks = [shard_range(M, r, world) for r in range(world)]
      assert blocks[0][0] == 0 and blocks[-1][1] == M
      assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
      sizes = [hi - lo for lo, hi in blocks]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    shard_range(8, 2, 2)


@pytest.mark.parametrize('world,M', [(2, 12), (3, 8)])
def test_sharded_ensemble_equals_single_process(tmp_path, world, M):
  import torch.multiprocessing as mp
  from emu.emu_backend import EmuBackend
  from pymoc_b200 import configs
  from pymoc_b200.ensemble import Ensemble
  nsteps = 49
  mp.spawn(_worker, args=(world, _free_port(), M, nsteps, str(tmp_path)), nprocs=world, join=True)
  got = np.load(os.path.join(str(tmp_path), 'gathered.npz'))
  spec = configs.c3_twocol_so(M, axes=(3, 2, 2, 1) if M == 12 else None)
  ens = Ensemble(spec, backend=EmuBackend())
  ens.run(nsteps)
  assert np.array_equal(got['b'], ens.state()['b_basin'])
  assert np.array_equal(got['psi'], ens.diagnostics()['Psi_iso_b'])


def test_sharded_spec_construction_equals_slice():
  """configs.members(lo, hi) builds exactly the members [lo, hi) of the global lattice."""
  from pymoc_b200 import configs
  for build, M in ((configs.c2_column_so, 64), (configs.c3_twocol_so, 64), (configs.c4_jansen_nadeau, 32),
                   (configs.twobasin, 8)):
    whole = build(M)
    lo, hi = M // 4 + 1, M - 3
    with configs.members(lo, hi):
      part = build(M)
    assert part.M == hi - lo
    for m in (0, (hi - lo) // 2, hi - lo - 1):
      a, b = part.member_case(m), whole.member_case(lo + m)

      def same(x, y):
        if isinstance(x, dict):
          return set(x) == set(y) and all(same(x[k], y[k]) for k in x)
        if x is None or y is None:
          return x is None and y is None
        return np.array_equal(np.asarray(x), np.asarray(y))
      assert same(a, b), (build.__name__, m)
