"""Writes tests/golden/*.json — the reference's own known-answer vectors for the scan+top-k path.

The numbers are transcribed from the reference test-suite (paths relative to /root/reference);
nothing here is computed by our code, so the fixtures pin the oracle, not the other way round.
Run:  python oracle/make_golden.py      (idempotent; output is committed)
"""
import json
import os

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# c/tests/neighbors/ann_cagra_c.cu:32-50 (also examples/c/src/{bruteforce,cagra}_c_example.c and
# java/cuvs-java/src/test/java/com/nvidia/cuvs/BruteForceAndSearchIT.java:60-75)
DATASET_4x2 = [[0.74021935, 0.9209938], [0.03902049, 0.9689629], [0.92514056, 0.4463501], [0.6673192, 0.10993068]]
QUERIES_4x2 = [[0.48216683, 0.0428398], [0.5084142, 0.6545497], [0.51260436, 0.2643005], [0.05198065, 0.5789965]]

golden = {
    "source": "rapidsai/cuvs @ 26.08.00 test-suite (see per-case 'ref')",
    "cases": [
        {
            "name": "cagra_c_4x2_k1", "ref": "c/tests/neighbors/ann_cagra_c.cu:32-50",
            "metric": "sqeuclidean", "k": 1, "dataset": DATASET_4x2, "queries": QUERIES_4x2,
            "neighbors": [[3], [0], [3], [1]],
            "distances": [[0.03878258], [0.12472608], [0.04776672], [0.15224178]],
            "eps": 1e-5,
        },
        {
            "name": "cagra_c_4x2_k1_bitset_filtered", "ref": "c/tests/neighbors/ann_cagra_c.cu:44-50",
            "metric": "sqeuclidean", "k": 1, "dataset": DATASET_4x2, "queries": QUERIES_4x2,
            "filter_keep": [0, 3],  # bitset 0b1001: ids 1 and 2 removed
            "neighbors": [[3], [0], [3], [0]],
            "distances": [[0.03878258], [0.12472608], [0.04776672], [0.59063464]],
            "eps": 1e-5,
        },
        {
            "name": "java_bruteforce_4x2_k3", "ref": "java/cuvs-java/src/test/java/com/nvidia/cuvs/BruteForceAndSearchIT.java:93-98",
            "metric": "sqeuclidean", "k": 3, "dataset": DATASET_4x2, "queries": QUERIES_4x2,
            "neighbors": [[3, 2, 0], [0, 2, 1], [3, 2, 0], [1, 0, 3]],
            "distances": [[0.038782537, 0.35904616, 0.83774555], [0.12472606, 0.21700788, 0.3191862],
                          [0.047766685, 0.20332813, 0.48305476], [0.15224183, 0.5906347, 0.5986643]],
            "eps": 1e-5,
        },
        {
            "name": "java_bruteforce_4x2_k3_filtered", "ref": "BruteForceAndSearchIT.java:114-126",
            "metric": "sqeuclidean", "k": 3, "dataset": DATASET_4x2, "queries": QUERIES_4x2,
            "filter_keep": [0, 1, 3],
            "neighbors": [[3, 0, 1], [0, 1, 3], [3, 0, 1], [1, 0, 3]],
            "distances": [[0.038782537, 0.83774555, 1.0540828], [0.12472606, 0.3191862, 0.32186073],
                          [0.047766685, 0.48305476, 0.7208309], [0.15224195, 0.5906347, 0.5986643]],
            "eps": 1e-5,
        },
    ],
    # cpp/tests/neighbors/brute_force.cu:169-184: k=2, L2Unexpanded; every neighbour of a point must
    # carry the point's own label (queries == dataset).
    "label_case": {
        "ref": "cpp/tests/neighbors/brute_force.cu:169-184", "k": 2, "metric": "l2_unexpanded",
        "points": [[2.7810836, 2.550537003], [1.465489372, 2.362125076], [3.396561688, 4.400293529],
                   [1.38807019, 1.850220317], [3.06407232, 3.005305973], [7.627531214, 2.759262235],
                   [5.332441248, 2.088626775], [6.922596716, 1.77106367], [8.675418651, -0.242068655],
                   [7.673756466, 3.508563011]],
        "labels": [0, 0, 0, 0, 0, 1, 1, 1, 1, 1],
    },
    # ivf_pq_fp_8bit.cuh:31-84 closed-form checks (ExpBits=5): kMin = 2^-15, kMax = 2^16*(2-1/8)
    "fp8": {
        "ref": "cpp/src/neighbors/ivf_pq/ivf_pq_fp_8bit.cuh:54-99",
        "unsigned": [[0.0, 0], [-1.0, 0], [1e-6, 0], [3.0517578125e-05, 0], [1.0, 120], [1.0625, 120],
                     [1.125, 121], [2.0, 128], [1e9, 255], [122880.0, 255], [122879.0, 254]],
        "decode_unsigned": [[120, 1.0625], [128, 2.125], [0, 3.24249267578125e-05]],
    },
}

os.makedirs(OUT, exist_ok=True)
with open(os.path.join(OUT, "reference_vectors.json"), "w") as f:
    json.dump(golden, f, indent=1)
print("wrote", os.path.join(OUT, "reference_vectors.json"))
