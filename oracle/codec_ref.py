"""TEST INFRASTRUCTURE ONLY (imported by tests/ alone): plain-Python restatement of the byte-wise 32-bit range
coder of contextgs_b200/csrc/entropy_codec.cu for STATIC-TABLE streams -- the counterpart of what torchac's
arithmetic coder does for the reference (utils/encodings.py:147-183; torchac itself is absent from the reference
tree, so its byte format cannot be pinned: parity for the codec is round-trip exactness).

Used to check, on the CPU and independently of the GPU decoder, that (a) GPU-encoded table streams decode to the
input symbols and (b) streams encoded here are byte-identical to the GPU's."""

TOP = 1 << 24
MASK32 = 0xFFFFFFFF


def encode(symbols, table_of, tables):
    """symbols: ints; table_of(i) -> table index of the i-th symbol; tables[t]: cumulative 16-bit frequencies."""
    low, rng, cache, cache_size, out = 0, MASK32, 0, 1, bytearray()
    first = True   # the byte cached at the start is always zero (the code value is < 1) and is not transmitted

    def shift_low():
        nonlocal low, cache, cache_size, first
        if (low & MASK32) < 0xFF000000 or (low >> 32) != 0:
            carry = (low >> 32) & 0xFF
            temp = cache
            while True:
                if first:
                    first = False
                else:
                    out.append((temp + carry) & 0xFF)
                temp = 0xFF
                cache_size -= 1
                if cache_size == 0:
                    break
            cache = (low >> 24) & 0xFF
        cache_size += 1
        low = (low & 0x00FFFFFF) << 8

    for i, s in enumerate(symbols):
        tb = tables[table_of(i)]
        lo, hi = int(tb[s]), int(tb[s + 1])
        r = rng >> 16
        low += r * lo
        rng = (r * (hi - lo)) & MASK32
        while rng < TOP:
            rng = (rng << 8) & MASK32
            shift_low()
    # termination: the smallest multiple of 2^24 that is >= low lies inside [low, low + range) (range >= 2^24); only its top
    # byte is written, the decoder reads zeros past the end of the stream
    low = (low + 0x00FFFFFF) & ~0x00FFFFFF
    shift_low()
    shift_low()
    return bytes(out)


def decode(data, n, table_of, tables):
    pos, code, rng = 0, 0, MASK32

    def nxt():
        nonlocal pos
        b = data[pos] if pos < len(data) else 0
        pos += 1
        return b

    for _ in range(4):
        code = ((code << 8) | nxt()) & MASK32
    out = []
    for i in range(n):
        tb = tables[table_of(i)]
        r = rng >> 16
        v = min(code // r, 0xFFFF)
        lo, hi = 0, len(tb) - 2
        while lo < hi:
            mid = lo + (hi - lo + 1) // 2
            if int(tb[mid]) <= v:
                lo = mid
            else:
                hi = mid - 1
        out.append(lo)
        code = (code - r * int(tb[lo])) & MASK32
        rng = (r * (int(tb[lo + 1]) - int(tb[lo]))) & MASK32
        while rng < TOP:
            code = ((code << 8) | nxt()) & MASK32
            rng = (rng << 8) & MASK32
    return out
