"""CPU oracle of the reference's forward-pass arithmetic -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this package.  The product (crane_b200/) never does and fails loudly without its
CUDA library.  Parity status per component is recorded in DESIGN.md section "Oracle".
"""
