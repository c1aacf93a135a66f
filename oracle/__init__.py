"""CPU oracle for the Eigenfaces hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy + one plain-C file), the arithmetic of the
reference's recognition and PCA-fit path (saladbkp/face-detection-recognization-PCA).
Every function cites the reference file:line it follows (paths are relative to
/root/reference, which exists only in the build container, never on the GPU box).

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
    legs may import or execute anything in here;
  * the product package (face-detection-recognization-pca_b200/) never imports it and has
    no CPU fallback: without the CUDA extension it raises.

Parity pinning status (see DESIGN.md section "Oracle"):
  * gen1 fit / projection: PINNED by the reference's shipped models/*.pkl and *_model_info.json
    (tests/golden/gen1_*.npz, produced by tests/golden/make_golden.py from /root/reference);
  * gen2 fit / recognition, cvtColor / resize integer spec: PINNED by outputs of the reference
    functions themselves (and of cv2 / sklearn, the un-vendored dependencies they call) run in the
    build container on seeded inputs and committed under tests/golden/;
  * L2 nearest neighbour and reconstruction error (north-star extras): PARITY UNPINNED -- the
    reference has no such code; the oracle is the textbook formula in float64.
"""
from . import preprocess, gen1, gen2, extras  # noqa: F401
