"""Reader/writer for the NGT1 named-tensor container used for fixtures and oracle dumps
(format defined in oracle/ref_hooks.c)."""
import gzip
import struct
import numpy as np


def read(path):
    out = {}
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        data = f.read()
    if data[:4] != b"NGT1":
        raise ValueError(f"{path}: not an NGT1 file")
    p = 4
    while p < len(data):
        (nl,) = struct.unpack_from("<H", data, p); p += 2
        name = data[p:p + nl].decode(); p += nl
        dtype = chr(data[p]); nd = data[p + 1]; p += 2
        dims = struct.unpack_from("<%dq" % nd, data, p); p += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        dt = np.float64 if dtype == "d" else np.int32
        nbytes = n * np.dtype(dt).itemsize
        if nd == 0:
            continue
        out[name] = np.frombuffer(data, dtype=dt, count=n, offset=p).reshape(dims).copy()
        p += nbytes
    return out


def write(path, arrays):
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(b"NGT1")
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            if a.dtype.kind == "f":
                a = a.astype(np.float64); code = b"d"
            else:
                a = a.astype(np.int32); code = b"i"
            nb = name.encode()
            f.write(struct.pack("<H", len(nb))); f.write(nb)
            f.write(code); f.write(struct.pack("<B", a.ndim))
            f.write(struct.pack("<%dq" % a.ndim, *a.shape))
            f.write(a.tobytes())


def scalar(d, key, default=None):
    if key not in d:
        return default
    return d[key].reshape(-1)[0].item()


# ---- LU pattern sets out of a recorded reference run (`*.trace.ngt.gz`, written by oracle/ref_hooks.c through
# tests/golden/make_golden.py): what Circuit.from_flat(lu_pattern=...) takes when a run has to follow KLU's own orders
def first_pattern(trace):
    ks = sorted({int(k.split("/")[0][1:]) for k in trace if k.endswith("/pat/n")})
    k = ks[0]
    pre = f"c{k}/pat/"
    return {kk[len(pre):]: v for kk, v in trace.items() if kk.startswith(pre)}


def run_patterns(trace):
    """the pivoting factors of a complete run, in the order the reference computed them (recorded by
    oracle/ref_hooks.c at every SMPreorder): the INITJCT iteration and the one after it, then the first two
    iterations of the first time point -- two factors for a UIC run.  Circuit.set_lu_pattern maps them onto
    pattern sets (identical factors share one)."""
    ks = sorted({int(k.split("/")[0][1:]) for k in trace if k.endswith("/pat/n")})
    uic = bool(int(trace[f"c{ks[0]}/mode"][0]) & 0x1000) or not (int(trace[f"c{ks[0]}/mode"][0]) & 0x200)
    want = 2 if uic else 4
    return [pattern_at(trace, k) for k in ks[:want]]


def pattern_at(trace, call):
    pre = f"c{call}/pat/"
    d = {kk[len(pre):]: v for kk, v in trace.items() if kk.startswith(pre)}
    return d or None
