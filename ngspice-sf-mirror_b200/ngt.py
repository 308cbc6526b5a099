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
