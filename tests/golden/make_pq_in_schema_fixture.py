"""Reads the reference's own binary fixture test_data/v0.27.1/pq_in_schema (written by Lance 0.27.1, used by
rust/lance/src/index/vector/ivf/v2.rs:2059 test_pq_storage_backwards_compat) with tools/lance_v2_reader.py and
stores its arrays as tests/golden/pq_in_schema.npz: the 512 x 32 f32 vectors, the IVF centroid, the PQ codebook, the
row ids and the TRANSPOSED `__pq_code` bytes exactly as merge_partitions wrote them.  Nothing here is computed by
our code.  Run in the build container (needs /root/reference):  python tests/golden/make_pq_in_schema_fixture.py
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools.lance_v2_reader import LanceV2File, read_ivf_pq_index  # noqa: E402

src = "/root/reference/test_data/v0.27.1/pq_in_schema"
data = LanceV2File(glob.glob(src + "/data/*.lance")[0])
ix = read_ivf_pq_index(glob.glob(src + "/_indices/*")[0])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pq_in_schema.npz")
np.savez_compressed(out, ids=data.column("id"), vectors=data.column("vec"), centroids=ix["centroids"],
                    lengths=ix["lengths"], codebook=ix["codebook"], codes_transposed=ix["codes_transposed"],
                    row_ids=ix["row_ids"], num_sub_vectors=ix["meta"]["num_sub_vectors"], nbits=ix["meta"]["nbits"],
                    transposed=ix["meta"]["transposed"], distance_type=ix["distance_type"])
print("wrote", out, os.path.getsize(out), "bytes")
