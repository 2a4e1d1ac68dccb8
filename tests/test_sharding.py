"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard a buffer by block
ranges, compress their shards (the oracle port stands in for the GPU codec -- test
infrastructure only), exchange sizes, and must reproduce the single-process stream."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from compression_algorithms_b200 import sharding  # noqa: E402

BLOCK = 65536


def test_block_ranges_cover_everything():
    for nblocks in (0, 1, 2, 7, 8, 9, 1526, 15259):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for r in range(world):
                first, cnt = sharding.block_range(nblocks, r, world)
                assert first == seen and cnt >= 0
                seen += cnt
            assert seen == nblocks
    assert sharding.block_range(15259, 7, 8) == (13356, 1903)      # 1 908 per GPU at G = 8
    assert sharding.byte_range(1_000_000_000, BLOCK, 7, 8) == (13356 * BLOCK, 1_000_000_000)
    assert sharding.byte_range(100, BLOCK, 1, 2) == (100, 100)    # fewer blocks than ranks: empty shard


def _oracle_fn(variant):
    from oracle import bindings as ob

    def fn(shard):
        a = shard.numpy()
        out, sizes = ob.port_lz77_compress_blocks(a, BLOCK, variant)
        nbytes = (sizes // 8 + 1) if variant == 0 else sizes
        stream = np.concatenate([out[b, : int(nbytes[b])] for b in range(len(sizes))])
        off = np.concatenate([[0], np.cumsum(nbytes)]).astype(np.int64)
        return torch.from_numpy(stream), torch.from_numpy(off), torch.from_numpy(sizes.astype(np.int64))
    return fn


def _worker(rank, world, port, n, variant, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from compression_algorithms_b200 import corpus
        data = torch.from_numpy(corpus.generate(n, corpus.ENWIK, 77).copy())
        a, b = sharding.byte_range(n, BLOCK, rank, world)
        sh = sharding.compress_sharded(data[a:b], n, BLOCK, _oracle_fn(variant))
        nblocks = (n + BLOCK - 1) // BLOCK
        index = sharding.exchange_block_index(sh.block_off, nblocks, torch.device("cpu"))
        whole = sharding.gather_stream(sh)
        q.put((rank, sh.shard_sizes.tolist(), sh.shard_off.tolist(), index.numpy(), whole.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("variant,n", [(1, 5 * BLOCK + 1234), (0, 3 * BLOCK), (1, 1000)])
def test_two_ranks_reproduce_single_process_stream(variant, n):
    from oracle import bindings as ob
    if not ob.have_port():
        pytest.skip("oracle port not built")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + variant * 7 + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, variant, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from compression_algorithms_b200 import corpus
    data = torch.from_numpy(corpus.generate(n, corpus.ENWIK, 77).copy())
    stream, off, _ = _oracle_fn(variant)(data)
    for rank, sizes, shard_off, index, whole in res:
        assert sum(sizes) == stream.numel() and shard_off[0] == 0 and shard_off[-1] == stream.numel()
        assert np.array_equal(index, off.numpy())
        assert np.array_equal(whole, stream.numpy())


# ---------------------------------------------------------------- whole-buffer Huffman over two ranks
def _np_splice(dst, dst_bit, src, nbits):
    """numpy statement of b200_huffman_splice_dev: MSB-first bits of src OR-ed into dst at dst_bit."""
    bits = np.unpackbits(src.numpy().view(np.uint32).astype(">u4").view(np.uint8))[:nbits]
    d = dst.numpy().view(np.uint32)
    img = np.unpackbits(d.astype(">u4").view(np.uint8))
    img[dst_bit: dst_bit + nbits] |= bits
    d[:] = np.packbits(img).view(">u4").astype(np.uint32)


def _huff_worker(rank, world, port, n, cut, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from compression_algorithms_b200 import corpus
        from oracle import bindings as ob
        data = corpus.generate(n, corpus.ENWIK, 78).copy()
        shard = torch.from_numpy(data[:cut] if rank == 0 else data[cut:])

        def hist_fn(t):
            return torch.from_numpy(np.bincount(t.numpy(), minlength=256).astype(np.int64))

        def encode_fn(t, freq):
            codes, lens, _ = ob.port_huffman_build(freq.numpy().astype(np.uint64))
            words, bits = ob.port_huffman_encode(t.numpy(), codes, lens)
            return torch.from_numpy(words.view(np.int32).copy()), bits

        sh = sharding.huffman_whole_compress_sharded(shard, hist_fn, encode_fn)
        whole, nbytes = sharding.huffman_whole_gather(sh, sh.stream, _np_splice)
        q.put((rank, sh.bit_off.tolist(), whole.numpy().view(np.uint32), nbytes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,cut", [(300_000, 123_457), (70_001, 70_001), (4096, 1)])
def test_two_ranks_whole_buffer_huffman(n, cut):
    """One tree over the data of both ranks (all-reduce of the histogram), shard streams spliced at bit
    granularity: the result is the whole-buffer huffman_compress stream."""
    from oracle import bindings as ob
    if not ob.have_port():
        pytest.skip("oracle port not built")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + n + cut) % 2000
    procs = [ctx.Process(target=_huff_worker, args=(r, 2, port, n, cut, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from compression_algorithms_b200 import corpus
    want = ob.port_huffman_compress(corpus.generate(n, corpus.ENWIK, 78).copy())
    for rank, bit_off, whole, nbytes in res:
        assert bit_off[-1] == want["word_idx"] * 32 + want["bit_idx"]
        assert nbytes == want["buffer_size"]
        assert np.array_equal(whole, want["words"])
