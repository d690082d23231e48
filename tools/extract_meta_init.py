#!/usr/bin/env python3
"""Extract the initial-value constants of the reference's shipped graph (models/NoiseFlow/ckpt/model.ckpt.best.meta)
into a small JSON fixture, WITHOUT TensorFlow: a minimal protobuf wire-format reader walks
MetaGraphDef.graph_def(2).node(1) and decodes Const nodes' `value` attr (AttrValue.tensor(8) -> TensorProto).

Run in the build container only (reads /root/reference); the output is committed:
    python tools/extract_meta_init.py /root/reference/models/NoiseFlow/ckpt/model.ckpt.best.meta tests/golden/meta_init_constants.json

What it pins: the Conv2d1x1 LU parameterisation is initialised from a random ORTHOGONAL matrix
(borealisflows/layers.py:95, matrix_param.py:100-140), and the graph stores that decomposition as constants
(P, the 6-vectors of L and U in `fill_triangular_inverse` order, log_S, sign_S).  Reassembling
P.L.(U + diag(sign_S.exp(log_S))) with OUR vector ordering must give an orthogonal matrix; any wrong ordering does not.
"""
import json
import struct
import sys

import numpy as np


def varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        if c < 0x80:
            return r, i
        s += 7


def fields(b):
    """Yield (field_number, wire_type, value) of one message; value = int or memoryview slice."""
    i, n = 0, len(b)
    while i < n:
        key, i = varint(b, i)
        f, w = key >> 3, key & 7
        if w == 0:
            v, i = varint(b, i)
        elif w == 1:
            v = bytes(b[i:i + 8]); i += 8
        elif w == 2:
            ln, i = varint(b, i)
            v = b[i:i + ln]; i += ln
        elif w == 5:
            v = bytes(b[i:i + 4]); i += 4
        else:
            raise ValueError("wire type %d" % w)
        yield f, w, v


DT = {1: ("<f4", 4), 2: ("<f8", 8), 3: ("<i4", 4), 9: ("<i8", 8)}


def tensor(b):
    dtype, shape, content, fvals, ivals, dvals = 0, [], None, [], [], []
    for f, w, v in fields(b):
        if f == 1:
            dtype = v
        elif f == 2:
            for f2, _, v2 in fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in fields(v2):
                        if f3 == 1:
                            size = v3
                    shape.append(size)
        elif f == 4:
            content = bytes(v)
        elif f == 5:
            fvals += list(np.frombuffer(bytes(v), "<f4")) if w == 2 else [struct.unpack("<f", v)[0]]
        elif f == 6:
            dvals += list(np.frombuffer(bytes(v), "<f8")) if w == 2 else [struct.unpack("<d", v)[0]]
        elif f == 7:
            if w == 2:
                j, vb = 0, bytes(v)
                while j < len(vb):
                    x, j = varint(vb, j)
                    ivals.append(x)
            else:
                ivals.append(v)
    if dtype not in DT:
        return None
    np_dt = DT[dtype][0]
    count = int(np.prod(shape)) if shape else 1
    if content is not None:
        arr = np.frombuffer(content, np_dt)
    else:
        ivals = [x - (1 << 64) if x >= (1 << 63) else x for x in ivals]   # varints carry negatives as 64-bit two's complement
        vals = {1: fvals, 2: dvals, 3: ivals, 9: ivals}[dtype]
        if not vals:
            vals = [0]
        arr = np.asarray(vals, np_dt)
        if arr.size == 1 and count > 1:
            arr = np.full(count, arr[0], np_dt)
    if arr.size != count:
        return None
    return arr.reshape(shape)


def main(meta_path, out_path):
    buf = memoryview(open(meta_path, "rb").read())
    graph = None
    for f, w, v in fields(buf):
        if f == 2 and w == 2:
            graph = v
    consts = {}
    for f, w, node in fields(graph):
        if f != 1:
            continue
        name = op = None
        value = None
        for f2, w2, v2 in fields(node):
            if f2 == 1:
                name = bytes(v2).decode()
            elif f2 == 2:
                op = bytes(v2).decode()
            elif f2 == 5:   # map<string, AttrValue> entry
                k = av = None
                for f3, _, v3 in fields(v2):
                    if f3 == 1:
                        k = bytes(v3).decode()
                    elif f3 == 2:
                        av = v3
                if k == "value" and av is not None:
                    for f4, _, v4 in fields(av):
                        if f4 == 8:
                            value = v4
        if op == "Const" and value is not None and (name.startswith("model/") or "Conv2d_1x1" in name) \
                and "/Adam" not in name and "gradients" not in name and "save" not in name:
            t = tensor(value)
            if t is not None and 0 < t.size <= 64:
                consts[name] = {"shape": list(t.shape), "dtype": str(t.dtype), "values": [float(x) for x in t.reshape(-1)]}
    keep = {}
    for k, v in consts.items():
        leaf = k.split("/")[-1]
        if ("Conv2d_1x1" in k and (leaf in ("initial_value", "input") or "Initializer" in k)) or "Initializer" in k:
            keep[k] = v
    json.dump({"source": "models/NoiseFlow/ckpt/model.ckpt.best.meta (reference, read without TensorFlow)",
               "script": "tools/extract_meta_init.py", "constants": keep}, open(out_path, "w"), indent=0, sort_keys=True)
    print("%d Const nodes under model/, %d kept -> %s" % (len(consts), len(keep), out_path))
    for k in sorted(keep)[:60]:
        print(k, keep[k]["shape"])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
