"""Multi-GPU parity (one process per GPU, NCCL): the sharded BPE trainer — whose per-merge delta reduction runs
inside the iteration-head kernel over peer memory — must produce the unsharded merge table, and the sharded
bounds (MIN/MAX all-reduce, gathered quantile) must equal the unsharded ones.  Skips on a one-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        sys.path.insert(0, ROOT)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from beast_tokenizer_b200 import BEASTBsplineTokenizer, FIGBPE
        from beast_tokenizer_b200.synth import SyntheticLoader, synth_device
        tok = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                                    gripper_indices=[6, 13], device=f"cuda:{rank}")
        tok.fit_parameters(SyntheticLoader(20, 32, 50, 14, seed0=1), verbose=False)      # local: never implicit
        n_chunks, per = 8, 5000
        chunk = lambda c: tok.encode(synth_device(per, 50, 14, 1000 + c, dev))[0]
        shard = torch.cat([chunk(c) for c in range(n_chunks) if c % world == rank])
        out = {"rank": rank}
        for mode in ("peer", "nccl"):
            os.environ["BEAST_B200_BPE_NO_PEER"] = "1" if mode == "nccl" else "0"
            fig = FIGBPE(vocab_size=1024, show_progress=False, device=f"cuda:{rank}")
            st = fig.fit_from_bins(shard)
            out[mode] = (st.tokenizer.merges_txt(), st.tokenizer.vocab_json(), st.min_token, st.max_token,
                         getattr(st.tokenizer, "trainer_mode", None))
        if rank == 0:
            full = torch.cat([chunk(c) for c in range(n_chunks)])
            ref = FIGBPE(vocab_size=1024, show_progress=False, device=f"cuda:{rank}", process_group=False).fit_from_bins(full)
            out["ref"] = (ref.tokenizer.merges_txt(), ref.tokenizer.vocab_json(), ref.min_token, ref.max_token)
        # uneven shards, one of them EMPTY (the last rank holds no sequence at all): same table as the unsharded run
        os.environ["BEAST_B200_BPE_NO_PEER"] = "0"
        if rank == world - 1:
            uneven = shard[:0]
        else:
            uneven = torch.cat([chunk(c) for c in range(n_chunks) if c % (world - 1) == rank])[: 3000 + 997 * rank]
        st = FIGBPE(vocab_size=700, show_progress=False, device=f"cuda:{rank}").fit_from_bins(uneven)
        out["uneven"] = (st.tokenizer.merges_txt(), st.tokenizer.vocab_json(), st.min_token, st.max_token)
        parts = [None] * world
        dist.all_gather_object(parts, uneven.cpu())
        if rank == 0:
            ref = FIGBPE(vocab_size=700, show_progress=False, device=f"cuda:{rank}", process_group=False).fit_from_bins(
                torch.cat(parts).to(dev))
            out["uneven_ref"] = (ref.tokenizer.merges_txt(), ref.tokenizer.vocab_json(), ref.min_token, ref.max_token)
        # sharded bounds: every rank passes its shard of the trajectories
        xs = [synth_device(3000, 50, 14, 77 + c, dev) for c in range(world)]
        tok.set_process_group("world")
        tok.update_weights_bounds(xs[rank])
        out["minmax"] = (tok.w_min.cpu().numpy(), tok.w_max.cpu().numpy())
        tok.fit_parameters([{"actions": xs[rank][i:i + 500]} for i in range(0, 3000, 500)], verbose=False)
        out["quant"] = (tok.w_min.cpu().numpy(), tok.w_max.cpu().numpy())
        if rank == 0:
            allx = torch.cat([x.to(dev) for x in xs])
            t2 = BEASTBsplineTokenizer(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, gripper_zero_order=True,
                                       gripper_indices=[6, 13], device=f"cuda:{rank}")
            t2.update_weights_bounds(allx)
            out["minmax_ref"] = (t2.w_min.cpu().numpy(), t2.w_max.cpu().numpy())
            t2.fit_parameters([{"actions": allx}], verbose=False)
            out["quant_ref"] = (t2.w_min.cpu().numpy(), t2.w_max.cpu().numpy())
        q.put(out)
        dist.barrier()
        dist.destroy_process_group()
    except Exception as exc:                                # pragma: no cover
        import traceback
        q.put({"rank": rank, "error": f"{exc}\n{traceback.format_exc()}"})


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_trainer_and_bounds_match_unsharded(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for r in results:
        assert "error" not in r, r.get("error")
    by_rank = {r["rank"]: r for r in results}
    ref = by_rank[0]["ref"]
    for r in results:
        assert r["peer"][:4] == ref, f"rank {r['rank']}: peer-fused sharded merge table differs from the unsharded one"
        assert r["nccl"][:4] == ref, f"rank {r['rank']}: NCCL-per-merge sharded merge table differs"
        assert "peer-fused" in (r["peer"][4] or ""), r["peer"][4]          # the NVLink path really ran
        assert r["uneven"] == by_rank[0]["uneven_ref"], f"rank {r['rank']}: uneven / empty shards give a different table"
        for key in ("minmax", "quant"):
            lo, hi = r[key]
            lo0, hi0 = by_rank[0][key + "_ref"]
            assert np.array_equal(lo, lo0) and np.array_equal(hi, hi0), key
