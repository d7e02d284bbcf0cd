"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference hot path (gasparian/CRNN-OCR-lite utils.py + Keras 2.2.2 / TF 1.8
semantics).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (crnn-ocr-lite_b200/) never does.

PARITY UNPINNED: the reference has no tests or golden vectors and its Keras/TF stack cannot run in this
image; the oracle is pinned by independent cross-checks only (tests/test_oracle_*.py).
"""
