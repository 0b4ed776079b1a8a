"""GPU, world_size 2 (needs two devices: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`; skipped on a one-GPU
box): config 4's data path -- sequences sharded over ranks, one handle per GPU, NCCL all-gather of every round's logits and tokens
through crane_b200_comm_init / crane_b200_decode_batch_gather -- against the CPU oracle run sequence by sequence."""
import numpy as np
import pytest
import torch

from conftest import rel_err
import crane_b200
from crane_b200 import synth
from oracle.qwen3 import Qwen3Oracle

pytestmark = pytest.mark.gpu

N_SEQ, ROUNDS, WORLD = 6, 3, 2
CFG = synth.TINY_QWEN3_UNTIED


def _prompt(i):
    return synth.synth_token_ids(20 + 7 * i, CFG["vocab_size"], f"mg{i}")


def _worker(rank, world, uid_q, out_q):
    try:
        torch.cuda.set_device(rank)
        if rank == 0:
            uid = crane_b200.comm_unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=120)
        w = dict(synth.synth_checkpoint(CFG))
        m = crane_b200.Qwen3Model(CFG, device=rank, max_seq_len=256, max_batch=4)
        m.load_checkpoint(w.items())
        m.comm_init(uid, rank, world)
        mine = crane_b200.shard_sequences(N_SEQ, world, rank)
        slots, toks = [], []
        for j, g in enumerate(mine):
            s = 0 if j == 0 else m.seq_create()
            m.seq_select(s)
            ids = _prompt(g)
            toks.append(m.forward_step_argmax(ids, 0))
            slots.append(s)
        rounds = []
        for r in range(ROUNDS):
            all_tok, all_lg = m.decode_batch_gather(slots, toks, want_logits=True)
            rounds.append((all_tok.copy(), all_lg))
            toks = [int(t) for t in all_tok[rank * len(mine):(rank + 1) * len(mine)]]
        out_q.put((rank, rounds, None))
        m.close()
    except Exception as e:  # surfaces in the parent instead of a silent timeout
        out_q.put((rank, None, repr(e)))


@pytest.mark.skipif(torch.cuda.device_count() < WORLD, reason="needs two GPUs")
def test_two_rank_batch_decode_all_gather_against_oracle():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, WORLD, uid_q, out_q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        rank, rounds, err = out_q.get(timeout=600)
        assert err is None, f"rank {rank}: {err}"
        res[rank] = rounds
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # the oracle, one sequence at a time
    w = dict(synth.synth_checkpoint(CFG))
    ref_tok = np.zeros((ROUNDS, N_SEQ), np.int64)
    ref_lg = np.zeros((ROUNDS, N_SEQ, CFG["vocab_size"]), np.float32)
    for g in range(N_SEQ):
        orc = Qwen3Oracle(CFG, w)
        ids = _prompt(g)
        tok = int(orc.forward(ids, 0).numpy().argmax())
        for r in range(ROUNDS):
            lg = orc.forward([tok], len(ids) + r).numpy().reshape(-1)
            ref_lg[r, g] = lg
            tok = int(lg.argmax())
            ref_tok[r, g] = tok
    worst = 0.0
    for r in range(ROUNDS):
        t0, l0 = res[0][r]
        t1, l1 = res[1][r]
        assert np.array_equal(t0, t1) and np.array_equal(l0, l1)          # every rank holds the same gathered batch
        assert list(t0) == list(ref_tok[r])                                # rank-major == batch order, greedy tokens exact
        worst = max(worst, max(rel_err(l0[g], ref_lg[r, g]) for g in range(N_SEQ)))
    print(f"2-rank batch decode, {N_SEQ} sequences x {ROUNDS} rounds: gathered logits rel max {worst:.2e}")
    assert worst < 1e-3
