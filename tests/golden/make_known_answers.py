"""Transcribes the reference's own known-answer unit tests for the IVF-PQ hot path into
tests/golden/reference_known_answers.json.  Inputs AND expected outputs are the literals that
appear in the reference test sources (file:line given per case, relative to /root/reference/rust);
nothing here is computed by our code.  Run:  python tests/golden/make_known_answers.py
"""
import json
import os

cases = []

# lance-linalg/src/distance/l2.rs:281-300  test_euclidean_distance
cases.append(dict(
    name="l2_euclidean_distance", ref="lance-linalg/src/distance/l2.rs:281-300", op="l2_batch", d=8,
    frm=[float(v) for v in range(2, 10)],
    to=[float(v) for r in (range(0, 8), range(1, 9), range(2, 10), range(3, 11)) for v in r],
    expect=[32.0, 8.0, 0.0, 8.0], exact=True))
# l2.rs:302-315 test_not_aligned (same numbers at an unaligned offset)
cases.append(dict(
    name="l2_not_aligned", ref="lance-linalg/src/distance/l2.rs:302-315", op="l2_batch", d=8,
    frm=[float(v) for v in range(0, 10)][2:],
    to=[float(v) for r in (range(0, 6), range(0, 8), range(1, 9), range(2, 10), range(3, 11)) for v in r][6:],
    expect=[32.0, 8.0, 0.0, 8.0], exact=True))
# l2.rs:317-324 test_odd_length_vector
cases.append(dict(
    name="l2_odd_length", ref="lance-linalg/src/distance/l2.rs:317-324", op="l2_batch", d=5,
    frm=[float(v) for v in range(2, 7)], to=[float(v) for v in range(0, 5)],
    expect=[20.0], exact=True))
# l2.rs:326-375 test_l2_distance_cases
values = [0.25335717, 0.24663818, 0.26330215, 0.14988247, 0.06042378, 0.21077952, 0.26687378,
          0.22145681, 0.18319066, 0.18688454, 0.05216244, 0.11470364, 0.10554603, 0.19964123,
          0.06387895, 0.18992095, 0.00123718, 0.13500804, 0.09516747, 0.19508345, 0.2582458,
          0.1211653, 0.21121833, 0.24809816, 0.04078768, 0.19586588, 0.16496408, 0.14766085,
          0.04898421, 0.14728612, 0.21263947, 0.16763233]
q = [0.18549609, 0.29954708, 0.28318876, 0.05424477, 0.093134984, 0.21580857, 0.2951282,
     0.19866848, 0.13868214, 0.19819534, 0.23271298, 0.047727287, 0.14394054, 0.023316395,
     0.18589257, 0.037315924, 0.07037327, 0.32609823, 0.07344752, 0.020155912, 0.18485495,
     0.32763934, 0.14296658, 0.04498596, 0.06254237, 0.24348071, 0.16009757, 0.053892266,
     0.05918874, 0.040363103, 0.19913352, 0.14545348]
cases.append(dict(name="l2_distance_cases", ref="lance-linalg/src/distance/l2.rs:326-375",
                  op="l2_batch", d=32, frm=q, to=values, expect=[0.31935784], exact=False,
                  rel=1.1920929e-07))  # assert_relative_eq! default max_relative = f32::EPSILON
# l2.rs:431-447 test_uint8_l2_edge_cases
cases.append(dict(name="l2_u8_zero", ref="lance-linalg/src/distance/l2.rs:433-435", op="l2_u8",
                  x=[0] * 2048, y=[0] * 2048, expect=0.0, exact=True))
cases.append(dict(name="l2_u8_max", ref="lance-linalg/src/distance/l2.rs:437-446", op="l2_u8",
                  x=[0] * 2048, y=[255] * 2048, expect=float(255 ** 2 * 2048), exact=True))
# cosine.rs:361-374 test_cosine (scipy / sklearn literals)
cases.append(dict(name="cosine_scipy", ref="lance-linalg/src/distance/cosine.rs:361-367",
                  op="cosine", x=[float(v) for v in range(1, 9)],
                  y=[float(v) for v in range(100, 108)], expect=1.0 - 0.900957, rel=1e-5))
cases.append(dict(name="cosine_sklearn", ref="lance-linalg/src/distance/cosine.rs:369-374",
                  op="cosine", x=[3.0, 45.0, 7.0, 2.0, 5.0, 20.0, 13.0, 12.0],
                  y=[2.0, 54.0, 13.0, 15.0, 22.0, 34.0, 50.0, 1.0], expect=1.0 - 0.87358063,
                  rel=1e-5))
# cosine.rs:386-393 test_cosine_not_aligned
cases.append(dict(name="cosine_not_aligned", ref="lance-linalg/src/distance/cosine.rs:386-393",
                  op="cosine", x=[16.0, 32.0], y=[1.0, 2.0], expect=0.0, abs=1e-6))
# pq/distance.rs:337-365 test_compute_on_transposed_codes: fully deterministic inputs; the
# reference asserts transposed scan == row-major scan.  We store the INPUT recipe; the expected
# relation (and the value of the first distances computed by hand below) is checked in the test.
cases.append(dict(name="pq_transposed_equals_rowmajor", ref="lance-index/src/vector/pq/distance.rs:337-365",
                  op="pq_scan_identity", num_vectors=100, num_sub_vectors=4, num_bits=8, dimension=16))

# lance-linalg/src/simd/dist_table.rs:179-217 test_sum_4bit_dist_table_basic: 32 vectors, code_len 2,
# the 32-byte code pattern repeated to n * code_len bytes, dist_table[i] = i % 16 + 1; the reference asserts
# kernel == scalar and dists[1] == 38
pattern = [0x12, 0x34, 0x56, 0x78, 0x9a, 0xbc, 0xde, 0xf0, 0x11, 0x22, 0x33, 0x44, 0x55, 0x66, 0x77, 0x88,
           0x99, 0xaa, 0xbb, 0xcc, 0xdd, 0xee, 0xff, 0x00, 0x12, 0x34, 0x56, 0x78, 0x9a, 0xbc, 0xde, 0xf0]
cases.append(dict(name="sum_4bit_dist_table_basic", ref="lance-linalg/src/simd/dist_table.rs:179-217",
                  op="sum_4bit_dist_table", n=32, code_len=2, codes=pattern * (32 * 2 // len(pattern)),
                  dist_table=[i % 16 + 1 for i in range(16 * 4)], expect_index=1, expect=38))

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_known_answers.json")
with open(out, "w") as f:
    json.dump(cases, f, indent=1)
print("wrote", out, len(cases), "cases")
