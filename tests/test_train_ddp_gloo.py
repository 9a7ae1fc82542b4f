"""Training mode (BASELINE config 5): one minibatch of Trainer.forward (reference
train_mono.py:159-194) around the fused loss, and its data-parallel form -- one process per rank,
DistributedDataParallel all-reduce of the NETWORK gradients as the only collective.

Runs in the GPU-less container: the operators are pointed at the test-only emulator build and the
ranks talk over gloo (world_size 2).  The property checked for DDP: the gradients every rank holds
after the all-reduce equal the mean of the single-process gradients of the two sub-batches -- each
rank's loss normalises by its own mask sums, exactly like a single-GPU run on that sub-batch would
(which is why they are NOT the gradients of the concatenated batch; see training.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from goldens import rel_l2


def _use_emulator():
    from emu_lib import emu
    from tcsfm_b200 import _cabi, ops
    ops.lib = emu
    ops._require_cuda = lambda *a: None
    ops.ARITH_FLAGS = _cabi.ARITH_CPU


def _frames(seed, b=2, h=24, w=40):
    from tcsfm_b200 import synth
    return synth.make_frames(b, h, w, seed=seed)


def _config(num_scales=1, full=True):
    from tcsfm_b200 import training
    return training.default_config(num_scales=num_scales, iterations=2, full_profile=full)


def _grads(step):
    return [p.grad.clone() for p in step.parameters() if p.grad is not None]


@pytest.fixture()
def emulated(monkeypatch):
    from emu_lib import emu
    from tcsfm_b200 import _cabi, ops
    monkeypatch.setattr(ops, "lib", emu)
    monkeypatch.setattr(ops, "_require_cuda", lambda *a: None)
    monkeypatch.setattr(ops, "ARITH_FLAGS", _cabi.ARITH_CPU)


@pytest.mark.parametrize("num_scales", [1, 2])
def test_train_step_matches_oracle_backend(emulated, num_scales):
    """The fused step (warp + stack per egomotion iteration, frame loss, smoothness, pose consistency) against the
    same step assembled from the oracle's restatement of the reference."""
    from oracle import ref_torch as O
    from tcsfm_b200 import training

    class OracleBackend:
        solve_pose_iteratively = staticmethod(O.iterative_pose)
        disp_to_depth = staticmethod(O.disp_to_depth)

        @staticmethod
        def compute_pose_consistency_loss(poses, poses_inv):            # train_mono.py:8-16
            total = 0
            for p, q in zip(poses, poses_inv):
                total += (p[:, 0:6] + q[:, 0:6]).abs()
            return total.mean()

        @staticmethod
        def make_loss(config):
            return lambda *a: O.compute_loss(config, *a)

    cfg = _config(num_scales)
    fr = _frames(5)
    res = []
    for backend in (training.Backend, OracleBackend):
        step, optim = training.make_step(cfg, seed=3, padded=False, backend=backend)
        total = training.run_train_step(step, optim, fr)
        res.append((float(total), _grads(step)))
    (a, ga), (b, gb) = res
    assert abs(a - b) <= 1e-5 * abs(b), (a, b)
    assert len(ga) == len(gb) and len(ga) >= 6
    for x, y in zip(ga, gb):
        assert rel_l2(x, y) < 2e-4, rel_l2(x, y)


def test_padded_networks_have_the_reference_parameter_volume():
    from tcsfm_b200 import training
    step, _ = training.make_step(_config(), padded=True)
    n = sum(p.numel() for p in step.parameters())
    assert n == training.REFERENCE_DEPTH_PARAMS + training.REFERENCE_POSE_PARAMS        # 15.69 M -> 62.8 MB of fp32 gradients


def _ddp_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    _use_emulator()
    from tcsfm_b200 import training
    dist.init_process_group("gloo", rank=rank, world_size=world)
    step, optim = training.make_step(_config(), seed=3, padded=False)
    model = training.wrap_ddp(step)
    training.run_train_step(model, optim, _frames(10 + rank))          # each rank: its own sub-batch
    grads = _grads(step)
    if rank == 0:
        torch.save(grads, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_two_ranks_average_the_sub_batch_gradients(tmp_path, emulated):
    from tcsfm_b200 import training
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "ddp_grads.pt")
    mp.spawn(_ddp_worker, args=(2, port, out_path), nprocs=2, join=True)
    ddp_grads = torch.load(out_path)
    singles = []
    for rank in range(2):
        step, optim = training.make_step(_config(), seed=3, padded=False)
        training.run_train_step(step, optim, _frames(10 + rank))
        singles.append(_grads(step))
    assert len(ddp_grads) == len(singles[0])
    for g, a, b in zip(ddp_grads, singles[0], singles[1]):
        want = 0.5 * (a + b)
        assert rel_l2(g, want) < 1e-5, rel_l2(g, want)


def test_flat_grad_trainer_matches_the_plain_step(emulated):
    """FlatGradTrainer (eager form of the CUDA-graphed step): same loss, same gradients, same Adam update as
    run_train_step -- the gradients live in one flat buffer."""
    from tcsfm_b200 import training
    cfg = _config()
    fr = _frames(5)
    step_a, optim_a = training.make_step(cfg, seed=3, padded=False)
    total_a = training.run_train_step(step_a, optim_a, fr)
    step_b, optim_b = training.make_step(cfg, seed=3, padded=False)
    trainer = training.FlatGradTrainer(step_b, optim_b, fr, use_graph=False)
    total_b = trainer.run(fr)
    assert torch.allclose(total_a, total_b, rtol=1e-6, atol=0)
    for p, q in zip(step_a.parameters(), step_b.parameters()):
        assert q.grad.data_ptr() >= trainer.flat.data_ptr() and q.grad.data_ptr() < trainer.flat.data_ptr() + 4 * trainer.flat.numel()
        if p.grad is not None:
            assert rel_l2(q.grad, p.grad) < 1e-6
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-9)
    # a second step on other frames reuses the static buffers
    total_b2 = trainer.run(_frames(6))
    total_a2 = training.run_train_step(step_a, optim_a, _frames(6))
    assert torch.allclose(total_a2, total_b2, rtol=1e-5, atol=0)


def _flat_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    _use_emulator()
    from tcsfm_b200 import training
    dist.init_process_group("gloo", rank=rank, world_size=world)
    step, optim = training.make_step(_config(), seed=3, padded=False)
    fr = _frames(10 + rank)
    trainer = training.FlatGradTrainer(step, optim, fr, group=dist.group.WORLD, use_graph=False)
    trainer.run(fr)
    if rank == 0:
        torch.save([p.grad.clone() for p in step.parameters()], out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_trainer_two_ranks_average_the_sub_batch_gradients(tmp_path, emulated):
    """The one collective of the graphed data-parallel step: a single all-reduce (mean) of the flat gradient buffer."""
    from tcsfm_b200 import training
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "flat_grads.pt")
    mp.spawn(_flat_worker, args=(2, port, out_path), nprocs=2, join=True)
    got = torch.load(out_path)
    singles = []
    for rank in range(2):
        step, optim = training.make_step(_config(), seed=3, padded=False)
        training.run_train_step(step, optim, _frames(10 + rank))
        singles.append([p.grad if p.grad is not None else torch.zeros_like(p) for p in step.parameters()])
    for g, a, b in zip(got, singles[0], singles[1]):
        want = 0.5 * (a + b)
        assert rel_l2(g, want) < 1e-5 or float((g - want).abs().max()) < 1e-12, rel_l2(g, want)
