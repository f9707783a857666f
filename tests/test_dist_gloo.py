"""world_size-2 gloo test of the N>1 host logic (scatter of packed inputs, gather of per-rank
outputs, length-balanced sharding).  The per-rank "forward" is a stand-in that tags rows with
rank and input content -- the model itself needs a GPU and is covered by the -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dissc_b200 import dist as ddist
    r, w, _ = ddist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    dev = torch.device("cpu")
    B_local, T = 3, 5
    sb = ddist.ShardedBatch(rank, world, dev)
    if rank == 0:
        g = torch.Generator().manual_seed(0)
        code = torch.randint(0, 100, (world * B_local, T), generator=g)
        f0 = torch.randn(world * B_local, T, generator=g)
        spkr = torch.arange(world * B_local)
        lengths = torch.randint(1, T + 1, (world * B_local,), generator=g, dtype=torch.int32)
        c, f, s, l = sb.scatter(code, f0, spkr, lengths, B_local, T)
    else:
        c, f, s, l = sb.scatter(None, None, None, None, B_local, T)
    # stand-in for the local forward: (B_local, 2) = [sum(code)+spkr, rank]
    y = torch.stack([c.sum(1).float() + s.float(), torch.full((B_local,), float(rank))], dim=1)
    out = sb.gather(y)
    if rank == 0:
        want0 = code.sum(1).float() + spkr.float()
        assert torch.equal(out[:, 0], want0)
        assert torch.equal(out[:, 1], torch.arange(world).repeat_interleave(B_local).float())
    else:
        assert out is None
    # int16 waveforms are gathered as raw bytes (NCCL has no 16-bit integer type) and come back as int16
    w = (torch.arange(B_local * 6, dtype=torch.int32).reshape(B_local, 6) - 7 + 1000 * rank).to(torch.int16)
    g16 = sb.gather(w)
    if rank == 0:
        assert g16.dtype == torch.int16 and g16.shape == (world * B_local, 6)
        for r in range(world):
            want = (torch.arange(B_local * 6, dtype=torch.int32).reshape(B_local, 6) - 7 + 1000 * r).to(torch.int16)
            assert torch.equal(g16[r * B_local:(r + 1) * B_local], want)
        ret.put("ok")
    else:
        assert g16 is None
    # overlapped pipeline (one packed scatter, double-buffered gather): 5 steps through 2 buffers, every step's gathered
    # result must be that step's inputs routed to the right rank and back
    hop = 4

    def fake_forward(c_, f_, s_, l_, out):   # "waveform" = f(code, f0 sign, spkr, lengths, rank)
        base = (c_.sum(1) + s_ + l_.to(torch.int64) + (f_ > 0).sum(1) + 1000 * rank).to(torch.int16)
        out.copy_(base.view(-1, 1).expand(-1, hop * T) + torch.arange(hop * T, dtype=torch.int16))

    pipe = ddist.ScatterGatherPipeline(rank, world, dev, B_local, T, hop, fake_forward)
    for step in range(5):
        if rank == 0:
            gs = torch.Generator().manual_seed(100 + step)
            code_s = torch.randint(0, 100, (world * B_local, T), generator=gs)
            f0_s = torch.randn(world * B_local, T, generator=gs)
            spkr_s = torch.randint(0, 50, (world * B_local,), generator=gs)
            len_s = torch.randint(1, T + 1, (world * B_local,), generator=gs, dtype=torch.int32)
            packed = ddist.pack_inputs(code_s, f0_s, spkr_s, len_s, world)
            assert packed.shape == (world, ddist.packed_layout(B_local, T)[1])
            c0, f0_0, s0, l0 = ddist.unpack_inputs(packed[1], B_local, T)      # round trip of rank 1's row
            assert torch.equal(c0, code_s[B_local:2 * B_local]) and torch.equal(f0_0, f0_s[B_local:2 * B_local])
            assert torch.equal(s0, spkr_s[B_local:2 * B_local]) and torch.equal(l0, len_s[B_local:2 * B_local])
        else:
            packed = None
        i = pipe.step(packed)
        pipe.wait(i)
        if rank == 0:
            got = pipe.gathered[i]
            assert got.shape == (world, B_local, hop * T) and got.dtype == torch.int16
            for r in range(world):
                sl = slice(r * B_local, (r + 1) * B_local)
                base = (code_s[sl].sum(1) + spkr_s[sl] + len_s[sl].to(torch.int64) + (f0_s[sl] > 0).sum(1) + 1000 * r)
                want = base.to(torch.int16).view(-1, 1) + torch.arange(hop * T, dtype=torch.int16)
                assert torch.equal(got[r], want), (step, r)
    pipe.flush()
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_gather_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == "ok"


def test_pipeline_single_rank_is_identity():
    """world == 1: no collective, the forward writes the result buffer itself."""
    from dissc_b200 import dist as ddist
    B, T, hop = 2, 3, 2
    code = torch.arange(B * T).reshape(B, T)
    packed = ddist.pack_inputs(code, torch.ones(B, T), torch.tensor([5, 6]), torch.tensor([3, 2], dtype=torch.int32), 1)

    def fwd(c, f, s, l, out):
        out.copy_((c.sum(1) + s).to(torch.int16).view(-1, 1).expand(-1, hop * T))

    pipe = ddist.ScatterGatherPipeline(0, 1, torch.device("cpu"), B, T, hop, fwd)
    i = pipe.step(packed)
    pipe.wait(i)
    assert pipe.gathered[i].shape == (1, B, hop * T)
    assert pipe.gathered[i][0, :, 0].tolist() == [0 + 1 + 2 + 5, 3 + 4 + 5 + 6]


def test_pack_batch_pads_and_records_lengths():
    from dissc_b200.dist import pack_batch
    codes = [torch.tensor([1, 2, 3]), torch.tensor([4])]
    f0s = [torch.tensor([0.1, 0.2, 0.3]), torch.tensor([0.4])]
    code, f0, spkr, lengths = pack_batch(codes, f0s, [7, 9], T=4)
    assert code.tolist() == [[1, 2, 3, 0], [4, 0, 0, 0]]
    assert lengths.tolist() == [3, 1] and spkr.tolist() == [7, 9]
    assert f0[0, :3].tolist() == [0.1, 0.2, 0.3] or torch.allclose(f0[0, :3], torch.tensor([0.1, 0.2, 0.3]))
