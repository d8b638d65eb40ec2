"""Census of the RLE v2 runs in one stripe of an ORC file (uncompressed): per integer stream, how many runs of
each kind and how many values they hold.  Used to decide where the integer kernels spend their time."""
import sys, collections
sys.path.insert(0, '.')
from oracle.orc_oracle import OracleFile

W = [1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,26,28,30,32,40,48,56,64]
KIND = {0: 'PRESENT', 1: 'DATA', 2: 'LENGTH', 3: 'DICT_DATA', 5: 'SECONDARY', 6: 'ROW_INDEX'}

def census(s):
    pos, out = 0, collections.Counter()
    vals = collections.Counter()
    n = len(s)
    while pos < n:
        h = s[pos]; k = h >> 6
        if k == 0:
            rl = (h & 7) + 3; wb = ((h >> 3) & 7) + 1; pos += 1 + wb; name = 'SR'
        else:
            rl = (((h & 1) << 8) | s[pos + 1]) + 1; code = (h >> 1) & 31
            if k == 1:
                pos += 2 + (rl * W[code] + 7) // 8; name = 'DIR>=96' if rl >= 96 else ('DIR<=10' if rl <= 10 else 'DIRmid')
            elif k == 2:
                b3, b4 = s[pos + 2], s[pos + 3]
                t = W[b3 & 31] + ((b4 >> 5) & 7) + 1
                cfb = t if t <= 24 else 26 if t <= 26 else 28 if t <= 28 else 30 if t <= 30 else 32 if t <= 32 else (t + 7) // 8 * 8
                pos += 4 + ((b3 >> 5) & 7) + 1 + (rl * W[code] + 7) // 8 + ((b4 & 31) * cfb + 7) // 8; name = 'PATCHED'
            else:
                p = pos + 2
                for _ in range(2):
                    while s[p] & 0x80: p += 1
                    p += 1
                if code: p += ((rl - 2) * W[code] + 7) // 8
                pos = p; name = 'DELpacked' if code else ('DELfix<=10' if rl <= 10 else 'DELfix')
        out[name] += 1; vals[name] += rl
    return out, vals

f = OracleFile(open(sys.argv[1], 'rb').read())
st = f.stripes[0]
streams, enc, _ = f._stripe_footer(st)
names = ['<root>'] + [c for c in f.schema().names]
for s in streams:
    if s.kind in (0, 6) : continue
    t = f.types[s.column]
    raw = bytes(f.data[s.offset:s.offset + s.length])
    try:
        o, v = census(raw)
    except Exception as e:
        continue
    tot = sum(v.values())
    if tot == 0: continue
    print(f"col {s.column:2d} {names[s.column] if s.column < len(names) else '?':16s} {KIND.get(s.kind, s.kind):9s} bytes={s.length:9d} runs={sum(o.values()):7d} vals={tot:8d}  " +
          ' '.join(f"{k}:{o[k]}/{v[k]}" for k in sorted(o)))
