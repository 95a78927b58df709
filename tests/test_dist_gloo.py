"""CPU, world_size 2, gloo: the N > 1 plumbing (single arena broadcast, batch sharding, max-over-ranks timing)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from consistentid_b200 import dist as cdist
    r, l, w = cdist.init_from_env(backend="gloo")
    arena = torch.arange(1000, dtype=torch.float16) if r == 0 else torch.zeros(1000, dtype=torch.float16)
    cdist.broadcast_arena(arena, src=0)
    ok_bcast = bool(torch.equal(arena, torch.arange(1000, dtype=torch.float16)))
    s, e = cdist.shard_batch(5, r, w)
    mx = cdist.max_over_ranks(float(10 + r), device="cpu")
    cdist.barrier()
    q.put((r, ok_bcast, (s, e), mx))
    dist.destroy_process_group()


def test_two_rank_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert [r[1] for r in res] == [True, True]                  # weights bit-identical after the single broadcast
    assert [r[2] for r in res] == [(0, 3), (3, 5)]              # contiguous, disjoint, covering shards
    assert [r[3] for r in res] == [11.0, 11.0]                  # max over ranks


def test_shard_edges():
    from consistentid_b200.dist import shard_batch
    assert shard_batch(0, 0, 4) == (0, 0)
    assert [shard_batch(32, r, 8) for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]
    assert shard_batch(3, 3, 4) == (3, 3)


def _denoise_worker(rank, world, port, q):
    """End-to-end N = 2 path on the CPU: rank-local engines (torch-emulated kernels), ONE arena broadcast, batch shards, no other collective."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from tests import emulated_ops
    emulated_ops.install_permanent()
    from consistentid_b200 import dist as cdist
    from consistentid_b200.arch import UNetSpec
    from consistentid_b200.pipeline import B200Denoiser
    from consistentid_b200.scheduler import B200Scheduler
    from consistentid_b200.unet import B200UNet
    from oracle import synth
    from oracle.unet_ref import tiny_config
    r, _, w = cdist.init_from_env(backend="gloo")
    cfg = tiny_config("sd15")
    ref = synth.build_ref_unet(cfg, seed=1234 + 77 * r, rank=16)              # rank 1 starts from DIFFERENT weights on purpose
    sd = {k: v for k, v in ref.state_dict().items() if ".processor." not in k}
    eng = B200UNet(UNetSpec.from_config(cfg), sd, synth.adapter_state_dict(ref), dtype=torch.float32, device="cpu", rank=16)
    before = eng.params.arena.clone()
    cdist.broadcast_arena(eng.params.arena, src=0)
    changed = not torch.equal(before, eng.params.arena)
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    lat = synth.synth_latents(2, cfg.sample_size, cfg.sample_size, seed=0)      # the GLOBAL batch; each rank takes its shard
    s, e = cdist.shard_batch(2, r, w)
    out = B200Denoiser(eng, B200Scheduler("ddim"), use_cuda_graph=False)(lat[s:e], null, aug, txt, num_inference_steps=2, guidance_scale=5.0,
                                                                            start_merge_step=0)
    cdist.barrier()
    q.put((r, changed, (s, e), out.float().cpu().numpy()))       # numpy: a tensor's shared-memory handle would die with this process
    dist.destroy_process_group()


def test_two_rank_sharded_denoise_matches_single_process_oracle():
    from oracle import synth
    from oracle.loop_ref import denoise_sd15
    from oracle.schedulers_ref import make_scheduler
    from oracle.unet_ref import tiny_config
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7) % 1000
    ps = [ctx.Process(target=_denoise_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted((q.get(timeout=300) for _ in ps), key=lambda t: t[0])
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert [r[1] for r in res] == [False, True]                  # rank 0 keeps its arena, rank 1's was overwritten by the broadcast
    assert [r[2] for r in res] == [(0, 1), (1, 2)]
    cfg = tiny_config("sd15")
    ref = synth.build_ref_unet(cfg, seed=1234, rank=16)         # rank 0's weights
    null, aug, txt = synth.synth_prompts(cfg.cross_attention_dim)
    lat = synth.synth_latents(2, cfg.sample_size, cfg.sample_size, seed=0)
    want = denoise_sd15(ref, make_scheduler("ddim"), lat, null, aug, txt, 2, guidance_scale=5.0, start_merge_step=0)
    got = torch.cat([torch.from_numpy(r[3]) for r in res], dim=0)
    assert (got - want).abs().max().item() <= 5e-4 * max(1.0, want.abs().max().item())
