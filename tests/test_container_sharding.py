"""Host logic either side of the hot path: the LACB container and the multi-rank index / payload
gather (world_size 2, gloo, CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lac_b200 import container, sharding


def _fake_streams(n, seed):
    rng = np.random.default_rng(seed)
    nbits = rng.integers(0, 300, n)
    nbits[:2] = [0, 8]
    ntok = rng.integers(0, 64, n)
    streams = [rng.integers(0, 256, (int(b) + 7) // 8, dtype=np.uint8).tobytes() for b in nbits]
    return streams, ntok, nbits


def test_container_roundtrip_and_validation():
    streams, ntok, nbits = _fake_streams(37, 1)
    blob = container.pack(streams, ntok, nbits, prec=48, vocab=32000, chunk_tokens=2048)
    c = container.unpack(blob)
    assert (c.prec, c.vocab, c.chunk_tokens, c.quantiser) == (48, 32000, 2048, container.QUANT_LQ32)
    assert blob == container.pack_payload(b"".join(streams), ntok, nbits, 48, 32000, 2048)
    c2 = container.unpack(container.pack(streams, ntok, nbits, 48, 32000, 2048, batch_streams=256, tag=77))
    assert (c2.batch_streams, c2.tag) == (256, 77)
    assert c.streams() == streams
    assert np.array_equal(c.ntok, ntok) and np.array_equal(c.nbits, nbits)
    assert container.unpack(container.pack([], [], [], 48, 5, 16)).n_chunks == 0
    with pytest.raises(ValueError):
        container.unpack(blob[:-1])
    with pytest.raises(ValueError):
        container.unpack(b"XXXX" + blob[4:])
    with pytest.raises(ValueError):
        container.pack([b"ab"], [1], [3], 48, 5, 16)
    with pytest.raises(ValueError):  # a round-1 (version 1) file: its quantiser no longer exists
        container.unpack(blob[:4] + b"\x01\x00" + blob[6:])


def test_batch_spans_cover_whole_batches():
    for n_chunks in (0, 1, 5, 256, 257, 2048, 2049):
        for B in (1, 4, 256):
            for world in (1, 2, 3, 8):
                spans = sharding.batch_spans(n_chunks, B, world)
                assert spans[0][0] == 0 and spans[-1][1] == n_chunks
                assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
                assert all(b % B == 0 for b, _ in spans if b < n_chunks)


def test_chunk_range_partitions():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [sharding.chunk_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_chunks, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    streams, ntok, nbits = _fake_streams(n_chunks, 5)
    b, e = sharding.chunk_range(n_chunks, rank, world)
    g_ntok, g_nbits = sharding.gather_index(torch.from_numpy(ntok[b:e]), torch.from_numpy(nbits[b:e]), n_chunks)
    payload = sharding.gather_payload(sharding.concat_streams(streams[b:e], "cpu"), g_nbits, n_chunks)
    if rank == 0:
        blob = container.pack(container.Container(48, 2, 32000, 2048, g_ntok.numpy().astype(np.uint32),
                                                  g_nbits.numpy().astype(np.uint32), payload).streams(),
                              g_ntok.numpy(), g_nbits.numpy(), 48, 32000, 2048)
        q.put(blob)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_chunks", [5, 16])
def test_two_rank_gather_matches_single_process(n_chunks):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_chunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    blob = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    streams, ntok, nbits = _fake_streams(n_chunks, 5)
    assert blob == container.pack(streams, ntok, nbits, 48, 32000, 2048)


# ------------------------------------------------------------------ the sharded job with a stand-in coder
# compress_sharded / decompress_sharded are host logic around encode_batch / decode_batch callbacks (on the GPU: the
# model + coder step engine).  Here the callbacks are a trivial reversible stand-in, so the batching (full-shape
# batches, padding), the rank layout and the two gathers are exercised on CPU with gloo.
def _fake_encode(tokens, ntok):
    streams, nbits = [], []
    for row, n in zip(tokens, ntok):
        data = np.asarray(row[:n], dtype="<i4").tobytes()
        streams.append(data)
        nbits.append(8 * len(data))
    return streams, nbits


def _fake_decode_factory(chunk_tokens):
    def dec(streams, ntok):
        out = np.zeros((len(streams), chunk_tokens), dtype=np.int32)
        for i, (s, n) in enumerate(zip(streams, ntok)):
            out[i, :n] = np.frombuffer(s, dtype="<i4")[:n]
        return out
    return dec


def _job_worker(rank, world, port, n_tokens, chunk, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    toks = np.random.default_rng(11).integers(0, 1000, n_tokens).astype(np.int32)
    blob = sharding.compress_sharded(toks, chunk, B, _fake_encode, 48, 1000, "cpu", tag=5)
    blobs = [blob]
    dist.broadcast_object_list(blobs, src=0)
    back = sharding.decompress_sharded(blobs[0], _fake_decode_factory(chunk), "cpu")
    if rank == 0:
        q.put((blobs[0], back))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_tokens,chunk,B", [(1000, 16, 4), (37, 8, 3), (0, 8, 2), (64, 8, 16)])
def test_sharded_job_two_ranks_equals_one_rank(n_tokens, chunk, B):
    toks = np.random.default_rng(11).integers(0, 1000, n_tokens).astype(np.int32)
    single = sharding.compress_sharded(toks, chunk, B, _fake_encode, 48, 1000, "cpu", tag=5)
    assert np.array_equal(sharding.decompress_sharded(single, _fake_decode_factory(chunk), "cpu"), toks)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_job_worker, args=(r, 2, port, n_tokens, chunk, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    blob, back = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert blob == single                      # the file does not depend on the number of ranks
    assert np.array_equal(back, toks)
    c = container.unpack(blob)
    assert c.batch_streams == B and c.tag == 5
