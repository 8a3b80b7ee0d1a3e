"""Minimal .pgen writer for tests (plink-ng pgenlib on-disk format, hard calls only): mode 0x02 and mode 0x10 with
vrtypes 0 (raw), 1 (1-bit + difflist), 2/3 (LD-compressed +/- inversion), 4/6/7 (difflist vs constant)."""
import numpy as np


def _varint(v):
    out = bytearray()
    while v >= 128:
        out.append((v & 127) | 128); v >>= 7
    out.append(v)
    return bytes(out)


def _pack2(codes):
    codes = np.asarray(codes, dtype=np.uint8)
    pad = (-codes.size) % 4
    c = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8).tobytes()


def _difflist(sample_ct, ids, vals):
    ids = list(map(int, ids)); vals = list(map(int, vals))
    dl = len(ids)
    out = bytearray(_varint(dl))
    if dl == 0:
        return bytes(out)
    idb = 1 if sample_ct < 256 else 2 if sample_ct < 65536 else 3 if sample_ct < 16777216 else 4   # pgenlib BytesToRepresentNzU32(raw_sample_ct)
    ngroups = (dl + 63) // 64
    deltas = []
    for g in range(ngroups):
        out += int(ids[g * 64]).to_bytes(idb, "little")
        d = bytearray()
        for e in range(g * 64 + 1, min(dl, g * 64 + 64)):
            d += _varint(ids[e] - ids[e - 1])
        deltas.append(bytes(d))
    for g in range(ngroups - 1):
        out.append((len(deltas[g]) - 63) & 255)
    out += _pack2(vals)
    for d in deltas:
        out += d
    return bytes(out)


def encode_variant(codes, vrtype, base=None):
    codes = np.asarray(codes, dtype=np.uint8); n = codes.size
    if vrtype == 0:
        return _pack2(codes)
    if vrtype == 1:
        vals, cnt = np.unique(codes, return_counts=True)
        top = sorted(vals[np.argsort(-cnt)][:2].tolist()) if vals.size > 1 else [int(vals[0]), int(vals[0]) + 1 if vals[0] < 3 else 3]
        if len(top) < 2 or top[0] == top[1]:
            top = [min(top[0], 2), min(top[0], 2) + 1]
        lo, hi = top
        bits = np.zeros((n + 7) // 8 * 8, dtype=np.uint8); bits[:n] = (codes == hi)
        other = np.where((codes != lo) & (codes != hi))[0]
        return bytes([lo * 4 + (hi - lo)]) + np.packbits(bits, bitorder="little").tobytes() + _difflist(n, other, codes[other])
    if vrtype in (2, 3):
        tgt = codes.copy()
        if vrtype == 3:
            tgt = np.where(codes == 0, 2, np.where(codes == 2, 0, codes)).astype(np.uint8)
        diff = np.where(tgt != base)[0]
        return _difflist(n, diff, tgt[diff])
    if vrtype in (4, 6, 7):
        b = vrtype & 3
        diff = np.where(codes != b)[0]
        return _difflist(n, diff, codes[diff])
    raise ValueError(vrtype)


def _deltalist(sample_ct, ids):
    """A difflist without genotypes (the sample ids of a dosage list, vrtype bits 5-6 = 01)."""
    ids = list(map(int, ids)); dl = len(ids)
    out = bytearray(_varint(dl))
    if dl == 0:
        return bytes(out)
    idb = 1 if sample_ct < 256 else 2 if sample_ct < 65536 else 3 if sample_ct < 16777216 else 4
    ngroups = (dl + 63) // 64
    deltas = []
    for g in range(ngroups):
        out += int(ids[g * 64]).to_bytes(idb, "little")
        d = bytearray()
        for e in range(g * 64 + 1, min(dl, g * 64 + 64)):
            d += _varint(ids[e] - ids[e - 1])
        deltas.append(bytes(d))
    for g in range(ngroups - 1):
        out.append((len(deltas[g]) - 63) & 255)
    for d in deltas:
        out += d
    return bytes(out)


def encode_dosage(dos, dmode):
    """dos: (N,) uint16 with 65535 = no dosage for that sample.  dmode 1 = list, 2 = unconditional, 3 = bitarray (vrtype bits 5-6)."""
    dos = np.asarray(dos, dtype=np.uint16); n = dos.size
    have = np.where(dos != 65535)[0]
    if dmode == 2:
        return dos.astype("<u2").tobytes()
    vals = dos[have].astype("<u2").tobytes()
    if dmode == 1:
        return _deltalist(n, have) + vals
    if dmode == 3:
        bits = np.zeros((n + 7) // 8 * 8, dtype=np.uint8); bits[have] = 1
        return np.packbits(bits, bitorder="little").tobytes() + vals
    raise ValueError(dmode)


def write_pgen(codes, vrtypes=None, mode=0x10, dosages=None, dosage_modes=None):
    """codes: (M, N) uint8.  dosages: optional (M, N) uint16 (0..32768 = ALT dosage 0..2, 65535 = none for that sample) with
    dosage_modes[v] in {0 (no track), 1, 2, 3}: the records then carry a dosage track and 8-bit vrtypes.  Returns the file bytes."""
    codes = np.asarray(codes, dtype=np.uint8); M, N = codes.shape
    hdr = bytes([0x6c, 0x1b, mode]) + int(M).to_bytes(4, "little") + int(N).to_bytes(4, "little")
    if mode == 0x02:
        return hdr + bytes([0x40]) + b"".join(_pack2(codes[v]) for v in range(M))
    vrtypes = [0] * M if vrtypes is None else list(vrtypes)
    recs = []; base = None
    for v in range(M):
        t = vrtypes[v]
        if t in (2, 3) and base is None:
            t = 0
        rec = encode_variant(codes[v], t, base)
        if t not in (2, 3):
            base = codes[v].copy()
        if dosages is not None and dosage_modes is not None and dosage_modes[v]:
            rec = rec + encode_dosage(dosages[v], dosage_modes[v])
            t |= int(dosage_modes[v]) << 5
        recs.append((t, rec))
    maxlen = max(len(r) for _, r in recs)
    len_bytes = 1 if maxlen < 256 else 2 if maxlen < 65536 else 3
    wide = any(t > 15 for t, _ in recs)                  # dosage bits need 8-bit vrtypes: record-length encoding 4..7
    ctrl = (len_bytes - 1) | (4 if wide else 0) | 0x40
    nblk = (M + 65535) // 65536
    body = bytearray(); blocks = []
    for b in range(nblk):
        sub = recs[b * 65536:(b + 1) * 65536]
        if wide:
            tb = bytearray(t for t, _ in sub)
        else:
            tb = bytearray((len(sub) + 1) // 2)
            for i, (t, _) in enumerate(sub):
                tb[i >> 1] |= t << (4 * (i & 1))
        lb = b"".join(len(r).to_bytes(len_bytes, "little") for _, r in sub)
        blocks.append(bytes(tb) + lb)
    header_len = 12 + 8 * nblk + sum(len(x) for x in blocks)
    offs = []; pos = header_len
    for b in range(nblk):
        offs.append(pos)
        for _, r in recs[b * 65536:(b + 1) * 65536]:
            body += r; pos += len(r)
    out = hdr + bytes([ctrl]) + b"".join(int(o).to_bytes(8, "little") for o in offs) + b"".join(blocks) + bytes(body)
    return out
