"""Contiguous-slice sharding of the codec across the GPUs of one box (one process per GPU, torch.distributed).

The reference has no parallelism of any kind; this is the B200-side decomposition of
the same serial format, so that the ranks' byte ranges concatenated ARE the reference's single stream.

encode, rank r of W holding input slice x_r:
  1. K1 histogram of x_r; all_gather of the W local 256-bin histograms (2 KiB each)          -> global histogram
  2. every rank builds the identical code on the host (microseconds) and derives every rank's bit total
     B_q = sum_b len[b] * hist_q[b] from the gathered histograms: the exclusive scan S_r of the B_q is this
     rank's global start bit -- no second pass over the data and no second collective
  3. gh_encode(start_bit = S_r mod 128): the slice comes out already phase-aligned with the global stream,
     so no bulk shifting is ever needed; only the single byte straddling each shard boundary is exchanged
     (1-byte send to the left neighbour, OR-ed into its last byte)
  4. the last rank appends the end mark + 1-padding; rank 0 owns the header.
  Rank r then owns global payload bytes [ceil(S_r / 8), ceil(S_{r+1} / 8)).

decode, rank r holding its byte range of the payload:
  1. slices are re-cut at 16-byte boundaries of the global stream (only byte offsets are needed for that); in ONE
     neighbour exchange each rank fetches a 32-byte halo from its right neighbour and the 4 KiB that precede its
     slice from its left neighbour
  2. every rank first walks those 4 KiB from an arbitrary bit (a self-synchronising code has found the true
     codeword boundaries long before the end of them): where that walk leaves the halo is, almost surely, where the
     slice's first codeword starts. Then it self-synchronises its slice from there (gh_decode_sync)
  3. ONE all_gather of (exit_bit, n_symbols, eof, entry) per rank confirms that every rank's entry is its left
     neighbour's exit; only if one is not (never observed) the stale ranks re-synchronise and the gather repeats
  4. exclusive scan of the symbol counts -> each rank's output offset; gh_decode_write; output stays sharded.

Collectives are tiny (KiB) and latency-bound; bulk data never crosses NVLink."""
import numpy as np
import torch
import torch.distributed as dist

HALO = 32          # bytes fetched from the right neighbour (a codeword may straddle the slice end)
LEFT_HALO = 4096   # bytes before the slice start that are walked to find the slice's first codeword
HEAD = 8192        # room in front of the payload buffer for the left halo


def _ceil_div(a, b):
    return -(-a // b)


class ShardedCodec:
    def __init__(self, codec, group=None):
        self.codec = codec
        self.lib = codec.lib
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.dev = codec.device

    # ---- buffers that persist across steps ----------------------------------------------------------------
    def prepare(self, n_local_max):
        """allocate per-rank buffers once: worst-case payload for n_local_max input bytes"""
        cap = self.lib.compress_bound(n_local_max) + 256
        full = torch.empty(HEAD + cap, dtype=torch.uint8, device=self.dev)
        return {
            "payload_full": full,
            "payload": full[HEAD:],
            "out": torch.empty(n_local_max + 4096, dtype=torch.uint8, device=self.dev),
            "hists": torch.empty(self.world * 256, dtype=torch.int64, device=self.dev),
            "hist": torch.empty(256, dtype=torch.int64, device=self.dev),
            "enc_ws": torch.empty(self.lib.encode_workspace_bytes(n_local_max) + 256, dtype=torch.uint8, device=self.dev),
            "dec_ws": torch.empty(self.lib.decode_workspace_bytes(cap) + 512, dtype=torch.uint8, device=self.dev),
            "byte": torch.zeros(1, dtype=torch.uint8, device=self.dev),
            "halo": torch.zeros(HALO, dtype=torch.uint8, device=self.dev),
            "meta": torch.zeros(self.world * 4, dtype=torch.int64, device=self.dev),
            "mine": torch.zeros(4, dtype=torch.int64, device=self.dev),
            "mine_host": torch.zeros(4, dtype=torch.int64).pin_memory() if self.dev.type == "cuda" else torch.zeros(4, dtype=torch.int64),
        }

    # ---- encode -------------------------------------------------------------------------------------------
    def compress_shard(self, x, st):
        """x: this rank's contiguous input slice (uint8, on the device). Returns a dict describing this rank's
        byte range of the global payload (and, on rank 0, the header)."""
        c, lib, W, r = self.codec, self.lib, self.world, self.rank
        n = x.numel()
        c.histogram(x, out=st["hist"])
        dist.all_gather_into_tensor(st["hists"], st["hist"], group=self.group)
        hists = st["hists"].cpu().numpy().astype(np.uint64).reshape(W, 256)  # one D2H, also the sync point
        if (hists.sum(axis=1) == 0).any():  # decided from the gathered data: every rank raises together
            raise ValueError("sharded compress: a rank holds an empty slice")
        code = lib.build_code(hists.sum(axis=0))
        bits = [lib.payload_bits(code, hists[q], with_eof=(q == W - 1)) for q in range(W)]
        starts = np.concatenate([[0], np.cumsum(bits)]).astype(object)  # python ints: no overflow
        S, E = int(starts[r]), int(starts[r + 1])
        phase = S % 128
        base_byte = (S - phase) // 8  # global byte held by local byte 0 (a multiple of 16)
        payload = st["payload"]
        lib.encode(x.data_ptr(), n, code, payload.data_ptr(), payload.numel(), st["enc_ws"].data_ptr(),
                   st["enc_ws"].numel(), start_bit=phase, append_eof=(r == W - 1), stream=c._stream())
        # boundary bytes: a shard starting mid-byte hands its leading partial byte to the left neighbour
        ops = []
        if r > 0 and S % 8:
            ops.append(dist.P2POp(dist.isend, payload[S // 8 - base_byte: S // 8 - base_byte + 1], r - 1, group=self.group))
        if r < W - 1 and E % 8:
            ops.append(dist.P2POp(dist.irecv, st["byte"], r + 1, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if r < W - 1 and E % 8:
            idx = E // 8 - base_byte
            payload[idx:idx + 1].bitwise_or_(st["byte"])
        first = _ceil_div(S, 8) if r > 0 else 0
        last = _ceil_div(E, 8)
        total_bytes = _ceil_div(int(starts[W]), 8)
        header = lib.write_header(code) if r == 0 else None
        return {"code": code, "header": header, "payload": payload, "payload_full": st["payload_full"], "start_bit": S,
                "base_byte": base_byte, "first_byte": first,
                "end_byte": last, "payload_bytes": last - first, "total_bytes": total_bytes,
                "byte_starts": [(_ceil_div(int(starts[q]), 8) if q > 0 else 0) for q in range(W)] + [total_bytes]}

    # ---- decode -------------------------------------------------------------------------------------------
    def decompress_shard(self, enc, st):
        """enc: what compress_shard returned (only byte offsets, the payload bytes and the code are used -- not
        the bit offsets). Returns (output tensor view, n_symbols) for this rank's slice."""
        c, lib, W, r = self.codec, self.lib, self.world, self.rank
        code, payload = enc["code"], enc["payload"]
        byte_starts, total = enc["byte_starts"], enc["total_bytes"]
        # slice boundaries: owned-range starts rounded up to 16 bytes of the global stream
        cuts = [0] + [min(total, _ceil_div(byte_starts[q], 16) * 16) for q in range(1, W)] + [total]
        a, b = cuts[r], cuts[r + 1]
        if any(cuts[q + 1] - cuts[q] < 16 for q in range(W)):  # same data on every rank: they raise together
            raise ValueError("sharded decode: a rank's slice of the payload is shorter than 16 bytes")
        own = [byte_starts[q + 1] - byte_starts[q] for q in range(W)]
        # the left halo is used when every rank can serve it from its own bytes (else: entry guess 0 and rounds)
        use_left = all(own[q] >= LEFT_HALO + HALO and cuts[q + 1] - LEFT_HALO >= byte_starts[q] for q in range(W - 1))
        full, base = enc["payload_full"], enc["base_byte"] - HEAD  # full[i] holds global byte base + i
        # one neighbour exchange: right halo (the right neighbour's first HALO owned bytes, appended after our owned
        # range) and left halo (the bytes between cuts[r] - LEFT_HALO and our first owned byte)
        ops = []
        if r > 0:
            lo = byte_starts[r] - base
            ops.append(dist.P2POp(dist.isend, full[lo:lo + HALO], r - 1, group=self.group))
            if use_left:
                ops.append(dist.P2POp(dist.irecv, full[cuts[r] - LEFT_HALO - base: byte_starts[r] - base], r - 1, group=self.group))
        if r < W - 1:
            ops.append(dist.P2POp(dist.irecv, st["halo"], r + 1, group=self.group))
            if use_left:
                ops.append(dist.P2POp(dist.isend, full[cuts[r + 1] - LEFT_HALO - base: byte_starts[r + 1] - base], r + 1, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if r < W - 1:
            hi = enc["end_byte"] - base
            full[hi:hi + HALO].copy_(st["halo"])
        ptr = full.data_ptr() + (a - base)
        slice_bytes = b - a
        readable = slice_bytes + (8 if r < W - 1 else 0)
        ws = st["dec_ws"]
        wptr = ws.data_ptr() + (-ws.data_ptr()) % 256
        wbytes = ws.numel() - 256
        stream = c._stream()
        entry = 0  # the first-codeword position this rank currently assumes for its slice
        if use_left and r > 0:
            walk = lib.decode_sync(ptr - LEFT_HALO, LEFT_HALO, LEFT_HALO + 8, code, 0, True, wptr, wbytes, stream)
            entry = int(walk.exit_bit)
        res = lib.decode_sync(ptr, slice_bytes, readable, code, entry, True, wptr, wbytes, stream)
        rounds = 0
        mine_host = st["mine_host"]
        while True:
            mine_host[0], mine_host[1], mine_host[2], mine_host[3] = int(res.exit_bit), int(res.n_symbols), int(res.eof_found), entry
            st["mine"].copy_(mine_host, non_blocking=True)
            dist.all_gather_into_tensor(st["meta"], st["mine"], group=self.group)
            meta = st["meta"].cpu().numpy().reshape(W, 4)
            rounds += 1
            # every rank evaluates the same condition on the same gathered data
            want = [0] + [int(meta[q - 1][0]) for q in range(1, W)]
            stale = [want[q] != int(meta[q][3]) for q in range(W)]
            if not any(stale):
                break
            if stale[r]:
                entry = want[r]
                res = lib.decode_sync(ptr, slice_bytes, readable, code, entry, False, wptr, wbytes, stream)
        counts = [int(meta[q][1]) for q in range(W)]
        eofs = [int(meta[q][2]) for q in range(W)]
        first_eof = eofs.index(1) if 1 in eofs else W
        n_sym = counts[r] if r <= first_eof else 0
        out = st["out"]
        if n_sym > out.numel():
            raise RuntimeError("sharded decode: output buffer too small for this rank's slice")
        lib.decode_write(ptr, slice_bytes, readable, code, out.data_ptr(), n_sym, wptr, wbytes, stream)
        offset = sum(counts[q] for q in range(r) if q <= first_eof)
        self.last_decode = {"rounds": rounds, "offset": offset, "counts": counts, "first_eof": first_eof, "cuts": cuts,
                            "left_halo": bool(use_left)}
        return out[:n_sym], n_sym

    # ---- verification helper (not part of the codec path) -----------------------------------------------
    def verify_roundtrip(self, x, out, n_sym):
        """True on every rank iff the concatenation of the ranks' decoded slices equals the concatenation of
        the ranks' inputs. Decode slices are cut at 16-byte payload boundaries, so each rank's first few
        symbols live at the tail of its left neighbour's output: that tail is sent over before comparing."""
        W, r = self.world, self.rank
        sizes = torch.zeros(W * 2, dtype=torch.int64, device=self.dev)
        mine = torch.tensor([x.numel(), n_sym], dtype=torch.int64).to(self.dev)
        dist.all_gather_into_tensor(sizes, mine, group=self.group)
        sizes = sizes.cpu().numpy().reshape(W, 2)
        in_start = np.concatenate([[0], np.cumsum(sizes[:, 0])])
        out_start = np.concatenate([[0], np.cumsum(sizes[:, 1])])
        ok = bool(in_start[W] == out_start[W])
        ok = ok and all(out_start[q] >= in_start[q] and out_start[q + 1] >= in_start[q + 1] for q in range(W))
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64).to(self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if not int(flag.item()):
            return False
        head = int(out_start[r] - in_start[r])              # symbols of my input held by the left neighbour
        tail = int(out_start[r + 1] - in_start[r + 1])      # symbols I hold that belong to the right neighbour
        recv = torch.empty(max(head, 1), dtype=torch.uint8, device=self.dev)
        ops = []
        if r < W - 1 and tail:
            ops.append(dist.P2POp(dist.isend, out[n_sym - tail:n_sym].contiguous(), r + 1, group=self.group))
        if r > 0 and head:
            ops.append(dist.P2POp(dist.irecv, recv[:head], r - 1, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        mine_part = out[: n_sym - tail]
        good = torch.equal(recv[:head], x[:head]) if head else True
        good = good and torch.equal(mine_part, x[head:])
        flag = torch.tensor([1 if good else 0], dtype=torch.int64).to(self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(int(flag.item()))
