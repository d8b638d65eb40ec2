"""A lane-by-lane model, in plain Python, of the windowed Snappy decoder in orc_rust_b200/csrc/k_decompress.cu
(snappy_parse + snappy_copy): every "lane" parses the element that would start at its byte of a 32-byte window, four
rounds of pointer jumping find the real element starts, an exclusive scan places their output, bytes are produced per
output position with the owner found from the start-bit mask, back-references read the ring of recent output, and
copies that read their own window's output run afterwards in order.  The model asserts the invariants the kernel
relies on (the pointer-jumping chain equals the true chain, owners cover their bytes, no byte is read from the ring
after its slot was reused, nothing is read before it was written).  It is documentation and a test of the algorithm,
not of the kernel: tests/test_snappy_window_model.py runs it on the CPU, the GPU parity tests check the kernel."""

stats = dict(windows=0, elements=0, dep=0, blocks=0, far=0)
HIST=4096
ring_pos=[-1]*HIST
ring_val=[0]*HIST
def ring_put(a,b):
    ring_pos[a%HIST]=a; ring_val[a%HIST]=b
def ring_get(a):
    assert ring_pos[a%HIST]==a, ('ring miss',a)
    return ring_val[a%HIST]
def window(s, n, p, out, ulen):
    o = len(out)
    L = range(32)
    sp = s + bytes(64)
    hdr=[0]*32; ln=[0]*32; src=[0]*32; lit=[False]*32; adv=[0]*32; sane=[False]*32; nxt=[32]*32
    for lane in L:
        q = p + lane
        lo = int.from_bytes(sp[q:q+4], 'little'); b4 = sp[q+4]
        tag = lo & 0xff; t = tag & 3
        if t == 0:
            lit[lane] = True; l = tag >> 2; h = 1
            if l >= 60:
                extra = l - 59; h += extra
                raw = (lo >> 8) | (b4 << 24)
                l = raw if extra == 4 else raw & ((1 << (8*extra)) - 1)
            l = (l + 1) & 0xffffffff
            hdr[lane]=h; ln[lane]=l; src[lane]=q+h
        elif t == 1:
            hdr[lane]=2; ln[lane]=4+((tag>>2)&7); src[lane]=((tag>>5)<<8)|((lo>>8)&0xff)
        elif t == 2:
            hdr[lane]=3; ln[lane]=1+(tag>>2); src[lane]=(lo>>8)&0xffff
        else:
            hdr[lane]=5; ln[lane]=1+(tag>>2); src[lane]=(lo>>8)|(b4<<24)
        adv[lane] = hdr[lane] + (ln[lane] if lit[lane] else 0)
        inside = q < n
        sane[lane] = inside and ln[lane] != 0 and q + adv[lane] <= n
        nxt[lane] = min(lane + adv[lane], 32) if sane[lane] else 32
    reach = 1; jump = list(nxt)
    for r in range(4):
        add = 0
        for lane in L:
            if (reach >> lane) & 1 and jump[lane] < 32: add |= 1 << jump[lane]
        reach |= add
        j2 = [jump[jump[lane] & 31] for lane in L]
        jump = [j2[lane] if jump[lane] < 32 else 32 for lane in L]
    mine = [bool((reach >> lane) & 1) and (p + lane < n) for lane in L]
    # check against true chain
    chain = []; c = 0
    while c < 32 and p + c < n:
        chain.append(c)
        if not sane[c]: break
        c = c + adv[c]
    assert [l for l in L if mine[l]] == chain, (chain, [l for l in L if mine[l]])
    excl=[0]*32; acc=0
    for lane in L:
        excl[lane]=acc
        if mine[lane] and sane[lane]: acc += ln[lane]
    total = acc
    for lane in L:
        if mine[lane]:
            oo = o + excl[lane]
            if not sane[lane] or oo + ln[lane] > ulen or (not lit[lane] and (src[lane]==0 or src[lane] > oo)): return None
    dep = [mine[l] and not lit[l] and excl[l] + min(ln[l], src[l]) > src[l] for l in L]
    last = max(l for l in L if mine[l])
    long_lit = [lit[l] and ln[l] >= 128 for l in L]
    ranks = [l for l in L if mine[l]]
    W = [(excl[l], src[l], ln[l], lit[l], dep[l] or long_lit[l]) for l in ranks]
    long_len = ln[last] if long_lit[last] else 0
    body = total - long_len
    dw = bytearray(total)  # window output, relative
    done = bytearray(total)
    def rd(idx):  # idx relative to window start, may be negative
        if idx < 0: return out[o + idx]
        assert done[idx], "read of unwritten byte"
        return dw[idx]
    for v0 in range(0, body, 32):
        stats['blocks'] += 1
        starts = 0
        for l in L:
            if mine[l] and v0 <= excl[l] < v0 + 32: starts |= 1 << (excl[l]-v0)
        before = sum(1 for l in L if mine[l] and excl[l] < v0)
        vals = {}
        for lane in L:
            v = v0 + lane
            if v < body:
                r = before + bin(starts & (0xffffffff >> (31-lane))).count('1') - 1
                eo, sv, elen, elit, skip = W[r]
                assert eo <= v < eo + elen
                if not skip:
                    k = v - eo
                    if elit: vals[v] = s[sv + k]
                    else:
                        kk = k if k < sv else k % sv
                        idx = v - k - sv + kk
                        assert idx < 0
                        a = o + idx
                        if o + total - a <= HIST:
                            vals[v] = ring_get(a); assert vals[v] == out[a]
                        else:
                            stats['far'] += 1; vals[v] = out[a]
        for v, b in vals.items(): dw[v] = b; done[v] = 1; ring_put(o+v, b)
    for l in L:
        if dep[l]:
            stats['dep'] += 1
            for i in range(ln[l]):
                idx = excl[l] + i
                kk = i if i < src[l] else i % src[l]
                b = ring_get(o + excl[l] - src[l] + kk)
                assert b == rd(idx - src[l])
                dw[idx] = b; done[idx] = 1
            for i in range(ln[l]): ring_put(o + excl[l] + i, dw[excl[l]+i])
    if long_len:
        for i in range(long_len):
            dw[excl[last]+i] = s[src[last]+i]; done[excl[last]+i]=1
        for i in range(max(0, long_len-HIST), long_len): ring_put(o+excl[last]+i, s[src[last]+i])
    assert all(done)
    out += dw
    stats['windows'] += 1; stats['elements'] += len(ranks)
    return p + last + adv[last]

def decode(comp):
    s = bytes(comp); n = len(s); p = 0; ulen = 0; sh = 0
    while True:
        b = s[p]; p += 1; ulen |= (b & 0x7f) << sh; sh += 7
        if b < 0x80: break
    out = bytearray()
    for i in range(HIST): ring_pos[i]=-1
    while p < n:
        r = window(s, n, p, out, ulen)
        if r is None:
            raise RuntimeError("fallback on valid stream at %d" % p)
        p = r
    assert len(out) == ulen
    return bytes(out)

