"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference hot path (gasparian/CRNN-OCR-lite utils.py + Keras 2.2.2 / TF 1.8
semantics).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (crnn-ocr-lite_b200/) never does.

PARITY: the reference has no tests and its Keras/TF stack cannot run in this image.
  * forward + decode: PINNED against the reference's own input -> prediction pairs -- the seven mjsynth examples of its
    README figures (imgs/STN_examples/*.png: network input, spatial-transformer output, predicted label) run through its
    shipped weights reproduce the labels the reference predicted, mistakes included (tests/golden/reference_examples.npz,
    tests/test_golden.py); caveat: the figures are intensity-autoscaled, the tests state the window they assume.
  * CTC loss / gradient, backward, optimiser: PARITY UNPINNED (no reference-run outputs exist) -- pinned by independent
    cross-checks only (tests/test_oracle_*.py).
"""
