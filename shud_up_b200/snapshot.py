"""Reader for the binary snapshots oracle/ref_driver.cpp writes (name[48], int32 dtype,
int64 n, payload) and for their .npz form under tests/golden/."""
import struct

import numpy as np


def read_bin(path):
    out = {}
    with open(path, "rb") as fh:
        while True:
            hdr = fh.read(48 + 4 + 8)
            if len(hdr) < 60:
                break
            name = hdr[:48].split(b"\0", 1)[0].decode()
            dtype, n = struct.unpack("<iq", hdr[48:])
            dt = np.float64 if dtype == 0 else np.int32
            out[name] = np.frombuffer(fh.read(n * np.dtype(dt).itemsize), dtype=dt).copy()
    return out


def load(path):
    if str(path).endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    return read_bin(path)
