#!/usr/bin/env python
"""Writes tests/golden/tf_upstream_ctc.json: the known-answer vectors of TensorFlow's own CTC unit tests
(tensorflow/python/kernel_tests/ctc_loss_op_test.py::testBasic and ctc_decoder_ops_test.py::testCTCGreedyDecoder /
testCTCDecoderBeamSearch, unchanged between TF 1.0 and 1.15 -- the reference pins tensorflow==1.8.0, Dockerfile:61).

RECALLED CONSTANTS: tensorflow is not vendored under /root/reference and not installed here, and there is no network, so these numbers
were typed from memory of the upstream test files, not copied from a checkout.  What makes them trustworthy anyway: they are 60 loss-
gradient entries + 2 losses + 2 beam scores with 6 significant digits each, and the C restatement (oracle/ctc_oracle.c), written
independently from SURVEY Appendix A, reproduces every one of them to the printed precision (tests/test_oracle_ctc.py) -- a mis-remembered
table could not agree with an independent implementation in all 64 numbers."""
import json
import os

ctc_loss = {
    "source": "tensorflow/python/kernel_tests/ctc_loss_op_test.py::testBasic (TF 1.x)",
    "note": "inputs = log(input_prob_matrix) fed as logits (depth 6, blank = 5, T = 5); loss = -loss_log_prob; gradient wrt the logits",
    "depth": 6, "seq_len": 5,
    "targets": [[0, 1, 2, 1, 0], [0, 1, 1, 0]],
    "loss": [3.34211, 5.42262],
    "input_prob_matrix": [
        [[0.633766, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
         [0.111121, 0.588392, 0.278779, 0.0055756, 0.00569609, 0.010436],
         [0.0357786, 0.633813, 0.321418, 0.00249248, 0.00272882, 0.0037688],
         [0.0663296, 0.643849, 0.280111, 0.00283995, 0.0035545, 0.00331533],
         [0.458235, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]],
        [[0.30176, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
         [0.24082, 0.397533, 0.0557226, 0.0546814, 0.0557528, 0.19549],
         [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, 0.202456],
         [0.280884, 0.429522, 0.0326593, 0.0339046, 0.0326856, 0.190345],
         [0.423286, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]]],
    "gradient": [
        [[-0.366234, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
         [0.111121, -0.411608, 0.278779, 0.0055756, 0.00569609, 0.010436],
         [0.0357786, 0.633813, -0.678582, 0.00249248, 0.00272882, 0.0037688],
         [0.0663296, -0.356151, 0.280111, 0.00283995, 0.0035545, 0.00331533],
         [-0.541765, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]],
        [[-0.69824, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
         [0.24082, -0.602467, 0.0557226, 0.0546814, 0.0557528, 0.19549],
         [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, -0.797544],
         [0.280884, -0.570478, 0.0326593, 0.0339046, 0.0326856, 0.190345],
         [-0.576714, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]]],
}
greedy = {
    "source": "tensorflow/python/kernel_tests/ctc_decoder_ops_test.py::testCTCGreedyDecoder (TF 1.x)",
    "note": "depth 4 (blank = 3), max_time 6, seq_len [4, 5], merge_repeated=True; log_prob = sum of -log(max prob) over the valid frames",
    "seq_len": [4, 5],
    "input_prob_matrix": [
        [[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.4, 0.6], [0.0, 0.0, 0.4, 0.6], [0.0, 0.9, 0.1, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]],
        [[0.1, 0.9, 0.0, 0.0], [0.0, 0.9, 0.1, 0.0], [0.0, 0.0, 0.1, 0.9], [0.0, 0.9, 0.1, 0.1], [0.9, 0.1, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]]],
    "decoded": [[0, 1], [1, 1, 0]],
    "neg_log_prob_factors": [[1.0, 0.6, 0.6, 0.9], [0.9, 0.9, 0.9, 0.9, 0.9]],
}
beam = {
    "source": "tensorflow/python/kernel_tests/ctc_decoder_ops_test.py::testCTCDecoderBeamSearch (TF 1.x)",
    "note": "depth 6 (blank = 5), seq_len 5 of 6 stored frames (the 6th is the test's 'random entry added in at time=5'), inputs = log(p) + 2.0 "
            "(the decoder is invariant to the offset), beam_width=2, top_paths=2, merge_repeated=False; log_prob as the TF 1.x test lists them",
    "seq_len": 5, "beam_width": 2, "top_paths": 2, "merge_repeated": False,
    "input_prob_matrix": [[0.30999, 0.309938, 0.0679938, 0.0673362, 0.0708352, 0.173908],
                          [0.215136, 0.439699, 0.0370931, 0.0393967, 0.0381581, 0.230517],
                          [0.199959, 0.489485, 0.0233221, 0.0251417, 0.0233289, 0.238763],
                          [0.279611, 0.452966, 0.0204795, 0.0209126, 0.0194803, 0.20655],
                          [0.51286, 0.288951, 0.0243026, 0.0220788, 0.0219297, 0.129878],
                          [0.155251, 0.164444, 0.173517, 0.176138, 0.169979, 0.160671]],
    "decoded": [[1, 0], [0, 1, 0]],
    "log_prob": [0.584855, 0.389139],
}
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_upstream_ctc.json")
json.dump({"provenance": "recalled constants (see make_tf_upstream_golden.py)", "ctc_loss": ctc_loss, "greedy": greedy, "beam": beam},
          open(out, "w"), indent=1)
print(out)
