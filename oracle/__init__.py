"""CPU oracle for the ALIGNN edge-gated conv hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import, call or execute
anything in this package.  The product package (alignn_b200/) must never import
it; its CUDA path fails loudly when the extension is missing instead of falling
back to anything here.
"""
