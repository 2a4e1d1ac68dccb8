"""Multi-GPU sharding of the block-parallel codecs.

Blocks are independent (fresh table / own code table per block), so every rank owns a
contiguous range of blocks and compresses it with no data-path collective. The ONE exchange
of the path is an all-gather of the shard sizes (world x int64), from which every rank
derives the global byte offset of its shard in the final stream; an optional second
all-gather of the per-block sizes gives the global block index. Payload stays sharded
(decode needs nothing else) or is placed at its offset with `gather_stream`.

Whole-buffer Huffman (ONE tree for all the data, what the reference's huffman_compress does)
is the one mode with a second exchange: an all-reduce of the 256-bin histogram before the
table build, then an all-gather of the shard BIT counts; the shard streams are spliced
bit-granularly (`huffman_whole_compress_sharded`, `huffman_whole_gather`).

The functions take any torch.distributed backend: `nccl` on the GPUs, `gloo` in the CPU
tests (world_size 2), where `compress_fn` is supplied by the test.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist


def block_range(nblocks, rank, world):
    """(first_block, count) of `rank`: ceil(nblocks / world) blocks per rank, last ranks may be short or empty."""
    per = (nblocks + world - 1) // world
    first = min(rank * per, nblocks)
    return first, min(per, nblocks - first)


def byte_range(n, block_size, rank, world):
    """(start, end) byte range of `rank`'s blocks in an n-byte buffer."""
    nblocks = (n + block_size - 1) // block_size
    first, cnt = block_range(nblocks, rank, world)
    return min(first * block_size, n), min((first + cnt) * block_size, n)


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def exchange_sizes(local_bytes, device, group=None):
    """All-gather of the shard sizes. `local_bytes` is an int or a 1-element int64 tensor on
    `device` (stays on the device: no host sync on the GPU path).
    -> (sizes int64[world], offsets int64[world + 1]) on `device`."""
    rank, world = _world(group)
    mine = local_bytes if isinstance(local_bytes, torch.Tensor) else torch.tensor([int(local_bytes)], dtype=torch.int64, device=device)
    mine = mine.reshape(1).to(torch.int64)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    if world > 1:
        dist.all_gather_into_tensor(sizes, mine.contiguous(), group=group)
    else:
        sizes.copy_(mine)
    offsets = torch.zeros(world + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(sizes, 0)
    return sizes, offsets


def exchange_block_index(local_block_off, nblocks_global, device, group=None):
    """All-gather of the per-block sizes -> global int64[nblocks_global + 1] byte offsets of
    every block in the concatenated stream. local_block_off: int64[local_blocks + 1]."""
    rank, world = _world(group)
    per = (nblocks_global + world - 1) // world
    sizes = torch.zeros(per, dtype=torch.int64, device=device)
    lb = local_block_off.numel() - 1
    if lb > 0:
        sizes[:lb] = (local_block_off[1:] - local_block_off[:-1]).to(device)
    allsz = torch.zeros(world * per, dtype=torch.int64, device=device)
    if world > 1:
        dist.all_gather_into_tensor(allsz, sizes, group=group)
    else:
        allsz.copy_(sizes)
    # rank r's blocks sit at [r*per, r*per + count_r); ranks are contiguous, so dropping the
    # tail padding of the last non-empty ranks is a prefix cut
    keep = torch.cat([allsz[r * per: r * per + block_range(nblocks_global, r, world)[1]] for r in range(world)])
    off = torch.zeros(nblocks_global + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(keep, 0)
    return off


@dataclass
class ShardedStream:
    rank: int
    world: int
    n_global: int
    block_size: int
    byte_start: int            # this rank's input range
    byte_end: int
    stream: torch.Tensor       # uint8, this rank's compressed blocks back to back
    block_off: torch.Tensor    # int64[local_blocks + 1], local offsets into `stream`
    block_sizes: torch.Tensor  # int64[local_blocks] (bits for the standalone LZ77, bytes otherwise)
    shard_sizes: torch.Tensor  # int64[world]
    shard_off: torch.Tensor    # int64[world + 1] global offset of every shard


def compress_sharded(data_shard, n_global, block_size, compress_fn, group=None):
    """data_shard: this rank's bytes (byte_range(n_global, ...)). compress_fn(shard) ->
    (stream uint8, block_off int64[nb+1], block_sizes int64[nb]) on the shard's device."""
    rank, world = _world(group)
    start, end = byte_range(n_global, block_size, rank, world)
    assert data_shard.numel() == end - start, "shard does not match byte_range()"
    if end > start:
        stream, block_off, block_sizes = compress_fn(data_shard)
    else:
        stream = torch.empty(0, dtype=torch.uint8, device=data_shard.device)
        block_off = torch.zeros(1, dtype=torch.int64, device=data_shard.device)
        block_sizes = torch.zeros(0, dtype=torch.int64, device=data_shard.device)
    sizes, off = exchange_sizes(block_off[-1:], data_shard.device, group)
    return ShardedStream(rank, world, n_global, block_size, start, end, stream, block_off, block_sizes, sizes, off)


def gather_stream(sh, group=None):
    """Place every shard at its global offset: all ranks end up with the whole stream (an
    all-gather of the payload, padded to the largest shard). Meant for writing the final file
    or for tests; decoding works on the shards as they are."""
    total = int(sh.shard_off[-1].item())
    if sh.world == 1:
        return sh.stream[:total].clone()
    mx = int(sh.shard_sizes.max().item())
    mine = torch.zeros(mx, dtype=torch.uint8, device=sh.stream.device)
    mine[: int(sh.shard_sizes[sh.rank].item())] = sh.stream[: int(sh.shard_sizes[sh.rank].item())]
    allp = torch.empty(sh.world * mx, dtype=torch.uint8, device=sh.stream.device)
    dist.all_gather_into_tensor(allp, mine, group=group)
    out = torch.empty(total, dtype=torch.uint8, device=sh.stream.device)
    for r in range(sh.world):
        a, b = int(sh.shard_off[r].item()), int(sh.shard_off[r + 1].item())
        out[a:b] = allp[r * mx: r * mx + (b - a)]
    return out


def lz77_compress_fn(ctx, variant, block_size):
    """compress_fn for the GPU path: the deflate / standalone LZ77 of compression_algorithms_b200.device"""
    from . import device as dv

    def fn(shard):
        st = dv.lz77_encode(ctx, shard, variant, block_size)
        return st.out[: st.total_bytes], st.block_off, st.block_sizes
    return fn


def lz77_decompress_sharded(ctx, sh, variant):
    """Every rank decodes its own shard (no collective). -> uint8 tensor of byte_end - byte_start bytes."""
    from . import device as dv
    n = sh.byte_end - sh.byte_start
    if n == 0:
        return torch.empty(0, dtype=torch.uint8, device=sh.stream.device)
    st = dv.Lz77Stream(variant=variant, out=sh.stream, block_sizes=sh.block_sizes, block_off=sh.block_off, n=n,
                       block_size=sh.block_size, total_bytes=int(sh.block_off[-1].item()))
    return dv.lz77_decode(ctx, st)


# ------------------------------------------------------------------ whole-buffer Huffman over ranks
@dataclass
class HuffmanShardedStream:
    rank: int
    world: int
    stream: object             # device.HuffmanStream of this rank's shard (own words from bit 0, own decode index)
    bits: int                  # bits of this rank's stream
    freq: torch.Tensor         # int64[256] global histogram (identical on every rank)
    shard_bits: torch.Tensor   # int64[world]
    bit_off: torch.Tensor      # int64[world + 1] exclusive prefix: where every shard starts in the whole stream


def reduce_histogram(freq64, group=None):
    """Sum of the shard histograms over the ranks (in place; the all-reduce of SURVEY.md §8e)."""
    rank, world = _world(group)
    if world > 1:
        dist.all_reduce(freq64, op=dist.ReduceOp.SUM, group=group)
    return freq64


def whole_stream_size(total_bits):
    """(u32 words, buffer_size in bytes) of a whole stream of `total_bits` bits: the reference's
    `4*word_idx + ceil(bit_idx/8)` (/root/reference/algorithms/huffman/huffman.c:318-320)."""
    word_idx, bit_idx = divmod(int(total_bits), 32)
    return word_idx + (1 if bit_idx else 0), 4 * word_idx + (bit_idx + 7) // 8


def huffman_whole_compress_sharded(data_shard, hist_fn, encode_fn, group=None):
    """One Huffman table for the data of all ranks. hist_fn(shard) -> int64[256] on the shard's device;
    encode_fn(shard, freq64) -> (stream, bits). On the GPUs these are device.huffman_histogram and
    device.huffman_encode_with_freq (see `huffman_fns`); the gloo CPU tests supply their own."""
    rank, world = _world(group)
    freq = reduce_histogram(hist_fn(data_shard), group)
    stream, bits = encode_fn(data_shard, freq)
    sizes, off = exchange_sizes(int(bits), freq.device, group)
    return HuffmanShardedStream(rank, world, stream, int(bits), freq, sizes, off)


def huffman_whole_gather(sh, words, splice_fn, group=None):
    """All ranks end up with the whole-buffer word stream (what huffman_compress writes for the concatenated
    input): all-gather of the shard words padded to the largest shard, then every shard is spliced at its bit
    offset by splice_fn(dst_words, dst_bit, src_words, src_bits). words: this rank's int32 word tensor.
    -> (int32 words, buffer_size bytes)."""
    nwords, nbytes = whole_stream_size(int(sh.bit_off[-1].item()))
    out = torch.zeros(nwords + 1, dtype=torch.int32, device=words.device)
    mxw = (int(sh.shard_bits.max().item()) + 31) // 32
    mine = torch.zeros(max(mxw, 1), dtype=torch.int32, device=words.device)
    myw = (sh.bits + 31) // 32
    mine[:myw] = words[:myw]
    if sh.world > 1:
        allw = torch.empty(sh.world * mine.numel(), dtype=torch.int32, device=words.device)
        dist.all_gather_into_tensor(allw, mine, group=group)
    else:
        allw = mine
    for r in range(sh.world):
        b = int(sh.shard_bits[r].item())
        if b:
            splice_fn(out, int(sh.bit_off[r].item()), allw[r * mine.numel(): (r + 1) * mine.numel()], b)
    return out[:nwords], nbytes


def huffman_fns(ctx):
    """(hist_fn, encode_fn, splice_fn) of the GPU path for the two functions above."""
    from . import device as dv
    return (lambda shard: dv.huffman_histogram(ctx, shard),
            lambda shard, freq: dv.huffman_encode_with_freq(ctx, shard, freq),
            lambda dst, bit, src, nbits: dv.huffman_splice(ctx, dst, bit, src, nbits))
