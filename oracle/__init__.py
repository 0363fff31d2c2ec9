"""CPU oracle of the retrieval hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (image_recommender_b200/, main/) never does.

PARITY UNPINNED for the search arithmetic (faiss_cpu==1.10.0 is an un-vendored, absent
dependency and the reference holds no golden vectors for it); the build half is pinned
against the real reference code by tests/golden/make_golden.py.  See b2k_oracle.c.
"""
from .b2k_oracle import (  # noqa: F401
    build_oracle, bf16_rne, bf16_to_f32, dims_total, normalize_l2, pack, search_exact,
    scores_f64, merge_topk, synth_rows, synth_queries, synth_query_source, sumsq32, num_threads,
)
