"""Minimal reader for Lance v2.0 files (TEST INFRASTRUCTURE; not part of the product).

Just enough of the format (protos/file2.proto, rust/lance-file/src/v2/reader.rs:footer layout) to open the index
files the reference writes for an IVF_PQ index -- `index.idx` (IVF model: pb IVF message with the centroids
tensor, protos/index.proto) and `auxiliary.idx` (merge_partitions, rust/lance/src/index/vector/builder.rs:
938-1079: columns `_rowid`, `__pq_code` in TRANSPOSED layout + `storage_metadata` JSON with the PQ codebook,
lance-index/src/vector/pq/storage.rs:52-67) -- and plain data files: flat, uncompressed, non-null pages only.

    footer (last 40 bytes): u64 column-metadata-0 offset, u64 column-metadata-offset table, u64 global-buffer
    offset table, u32 #global buffers, u32 #columns, u16 major, u16 minor, "LANC"
    global buffer 0 = FileDescriptor { 1: Schema { 1: repeated Field{2 name, 5 logical_type}, 5: metadata map }, 2: rows }
    column metadata = ColumnMetadata { 2: repeated Page { 1 buffer_offsets, 2 buffer_sizes, 3 length } }
"""
import json
import struct

import numpy as np


def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, i


def pb_fields(b):
    """wire-level protobuf decode -> [(field number, wire type, value)]"""
    out, i = [], 0
    while i < len(b):
        key, i = _varint(b, i)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 2:
            ln, i = _varint(b, i)
            v = b[i:i + ln]
            i += ln
        elif wt == 5:
            v = struct.unpack("<I", b[i:i + 4])[0]
            i += 4
        elif wt == 1:
            v = struct.unpack("<Q", b[i:i + 8])[0]
            i += 8
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.append((f, wt, v))
    return out


def _packed_u64(v):
    out, i = [], 0
    while i < len(v):
        x, i = _varint(v, i)
        out.append(x)
    return out


def pb_tensor(b):
    """protos/index.proto Tensor {1 data_type, 2 shape, 3 data} -> numpy array"""
    dt, shape, data = 0, [], b""
    for f, wt, v in pb_fields(b):
        if f == 1:
            dt = v
        elif f == 2:
            shape += _packed_u64(v) if wt == 2 else [v]
        elif f == 3:
            data = v
    np_dt = {1: np.float16, 2: np.float32, 3: np.float64, 4: np.uint8, 5: np.uint16, 6: np.uint32, 7: np.uint64}[dt]
    return np.frombuffer(data, dtype=np_dt).reshape(shape).copy()


_ELEM = {"int64": np.int64, "uint64": np.uint64, "uint8": np.uint8, "float": np.float32, "halffloat": np.float16,
         "double": np.float64, "int32": np.int32, "uint32": np.uint32}


class LanceV2File:
    def __init__(self, path):
        self.b = b = open(path, "rb").read()
        assert b[-4:] == b"LANC", "not a Lance file"
        cm0, cmo, gbo, ngb, ncol, major, minor = struct.unpack("<QQQIIHH", b[-40:-4])
        assert (major, minor) in ((0, 3), (2, 0)), f"unsupported Lance file version {major}.{minor}"
        self.global_buffers = []
        for g in range(ngb):
            off, sz = struct.unpack("<QQ", b[gbo + 16 * g:gbo + 16 * g + 16])
            self.global_buffers.append(b[off:off + sz])
        self.fields, self.metadata, self.num_rows = [], {}, 0
        for f, wt, v in pb_fields(self.global_buffers[0]):
            if f == 1:  # Schema
                for f2, wt2, v2 in pb_fields(v):
                    if f2 == 1:
                        fld = {x[0]: x[2] for x in pb_fields(v2)}
                        self.fields.append((fld[2].decode(), fld[5].decode()))
                    elif f2 == 5:
                        kv = {x[0]: x[2] for x in pb_fields(v2)}
                        self.metadata[kv[1].decode()] = kv.get(2, b"")
            elif f == 2:
                self.num_rows = v
        self.columns = []
        for c in range(ncol):
            off, sz = struct.unpack("<QQ", b[cmo + 16 * c:cmo + 16 * c + 16])
            pages = []
            for f, wt, v in pb_fields(b[off:off + sz]):
                if f == 2:
                    pg = {"offsets": [], "sizes": [], "length": 0}
                    for f2, wt2, v2 in pb_fields(v):
                        if f2 == 1:
                            pg["offsets"] += _packed_u64(v2) if wt2 == 2 else [v2]
                        elif f2 == 2:
                            pg["sizes"] += _packed_u64(v2) if wt2 == 2 else [v2]
                        elif f2 == 3:
                            pg["length"] = v2
                    pages.append(pg)
            self.columns.append(pages)

    def column(self, name):
        """flat, non-null column -> numpy array ([rows] or [rows][dim] for fixed_size_list)"""
        idx = [n for n, _ in self.fields].index(name)
        logical = self.fields[idx][1]
        dim = 1
        if logical.startswith("fixed_size_list:"):
            _, elem, dim = logical.split(":")
            dim = int(dim)
        else:
            elem = logical
        dt = np.dtype(_ELEM[elem])
        parts = []
        for pg in self.columns[idx]:
            assert len(pg["offsets"]) == 1, "only flat, non-null pages are supported"
            off, sz = pg["offsets"][0], pg["sizes"][0]
            assert sz == pg["length"] * dim * dt.itemsize, "page is not a plain flat buffer"
            parts.append(np.frombuffer(self.b[off:off + sz], dtype=dt))
        a = np.concatenate(parts) if parts else np.zeros(0, dt)
        return a.reshape(-1, dim).copy() if dim > 1 else a.copy()


def read_ivf(buf):
    """pb IVF message (protos/index.proto): {2 offsets, 3 lengths, 4 centroids_tensor}"""
    offsets, lengths, cent = [], [], None
    for f, wt, v in pb_fields(buf):
        if f == 2:
            offsets += _packed_u64(v) if wt == 2 else [v]
        elif f == 3:
            lengths += _packed_u64(v) if wt == 2 else [v]
        elif f == 4:
            cent = pb_tensor(v)
    return np.asarray(offsets, np.uint64), np.asarray(lengths, np.uint32), cent


def read_ivf_pq_index(index_dir):
    """-> dict(centroids [K][d], lengths [K], codebook [M][2^nbits][d/M], codes_transposed (as stored, the bytes of
    the `__pq_code` column), codes [n][M] row-major, row_ids [n], meta)"""
    import os
    idx = LanceV2File(os.path.join(index_dir, "index.idx"))
    aux = LanceV2File(os.path.join(index_dir, "auxiliary.idx"))
    _, _, centroids = read_ivf(idx.global_buffers[1])
    offsets, lengths, _ = read_ivf(aux.global_buffers[1])
    meta = json.loads(json.loads(aux.metadata["storage_metadata"].decode())[0])
    assert meta.get("codebook_position", 0) == 0 and meta["codebook_tensor"], "codebook kept in a global buffer: not handled"
    cb = pb_tensor(bytes(meta.pop("codebook_tensor")))              # [2^nbits][d]: row c = concat_m cb[m][c]?  no:
    M, nbits, d = meta["num_sub_vectors"], meta["nbits"], meta["dimension"]
    # FixedSizeList(value_length = d) over the flat [M][2^nbits][d/M] values (pq/utils.rs:59-76)
    codebook = cb.reshape(-1).reshape(M, 1 << nbits, d // M)
    stored = aux.column("__pq_code")                                # [n][M] as a LIST column ...
    row_ids = aux.column("_rowid")
    n = len(row_ids)
    codes = np.empty((n, M), np.uint8)
    flat = stored.reshape(-1)
    if meta.get("transposed", False):                               # ... whose bytes are [M][n_p] per partition
        pos = 0
        for ln in lengths.tolist():
            blk = flat[pos * M:(pos + ln) * M].reshape(M, ln)
            codes[pos:pos + ln] = blk.T
            pos += ln
    else:
        codes[:] = stored
    return dict(centroids=centroids, lengths=lengths, offsets=offsets, codebook=codebook, codes_transposed=flat.copy(),
                codes=codes, row_ids=row_ids, meta=meta, distance_type=aux.metadata.get("distance_type", b"l2").decode())
