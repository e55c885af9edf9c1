"""Data-parallel semantics on CPU (gloo, world_size 2): two ranks, each with half of the minibatch, must
reproduce the single-process result on the global batch -- losses, loss-history matching weights and every
post-step parameter (fp64 test double for the op-set, so the comparison is exact to ~1e-10).
Covers: global-batch noise slicing, identical peer draws, the scalar all-reduce feeding (sum m / numel)^2 and
w_match, coefficient scaling by 1/world and the flat-gradient all-reduce before the fused Adam."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import council_oracle as co
from common import load_golden, setup_case

CASE = 'm2f64_n4_b2'


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(gold, x_a, x_b, states, hp):
    from council_gan_b200.trainer_council import Council_Trainer
    from ops_torch import TorchOps
    from test_trainer_host_cpu import _randn32, load_states
    co.seed_all(hp['random_seed'])
    tr = Council_Trainer(hp, 'cpu', _ops=TorchOps('cpu', torch.float64))
    load_states(tr, states)
    co.seed_all(gold['rng_seed'])
    saved = torch.randn
    torch.randn = _randn32(torch.float32)
    try:
        tr.dis_update(x_a, x_b, hp)
        tr.dis_council_update(x_a, x_b, hp)
        tr.gen_update(x_a, x_b, hp, gold['iteration'])
    finally:
        torch.randn = saved
    out = {'dis': [float(v) for v in tr.loss_dis_total_s], 'disc': [float(v) for v in tr.loss_dis_council_total_s],
           'gen': [float(v) for v in tr.loss_gen_total_s], 'w_match': float(tr.w_match_a2b_conf)}
    tr.synchronize()  # the last family's all-reduce + Adam are deferred under data parallelism
    for name, net in tr._nets.items():
        out['p_' + name] = net.bank.data.clone()
    return out


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    gold = load_golden(CASE)
    hp, states, x_a, x_b = setup_case(gold)
    b = x_a.size(0) // world
    out = _run(gold, x_a[rank * b:(rank + 1) * b], x_b[rank * b:(rank + 1) * b], states, hp)
    if rank == 0:
        ret.update(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_ranks_equal_one_rank_global_batch():
    import sys
    torch.set_num_threads(4)
    gold = load_golden(CASE)
    hp, states, x_a, x_b = setup_case(gold)
    single = _run(gold, x_a, x_b, states, hp)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    ret = dict(ret)
    for k in ('dis', 'disc', 'gen'):
        for a, b in zip(single[k], ret[k]):
            assert abs(a - b) <= 1e-7 * abs(a), (k, a, b)
    assert abs(single['w_match'] - ret['w_match']) < 1e-9
    for k, v in single.items():
        if k.startswith('p_'):
            diff = (v - ret[k]).abs().max().item()
            assert diff < 1e-7, (k, diff)  # << lr = 1e-4: same update up to fp64 summation order


def _init_worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from council_gan_b200.trainer_council import Council_Trainer
    from ops_torch import TorchOps
    gold = load_golden(CASE)
    hp, states, x_a, x_b = setup_case(gold)
    co.seed_all(100 + rank)  # ranks seeded DIFFERENTLY (seed + rank is common practice)
    tr = Council_Trainer(hp, 'cpu', _ops=TorchOps('cpu'))
    sums = torch.tensor([float(net.bank.data.double().sum()) for net in tr._nets.values()] +
                        [float(tr._nets['gen_a2b'].frozen.data.double().sum())], dtype=torch.float64)
    got = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(got, sums)
    if rank == 0:
        ret['equal'] = bool(torch.equal(got[0], got[1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ranks_seeded_differently_start_from_rank0_parameters():
    """Only gradients are all-reduced during training, so the constructor broadcasts rank 0's parameter banks."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_init_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert dict(ret)['equal'] is True
