"""CPU oracle for the OptiSpeech hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch/numpy restatement of the reference algorithm (mush42/optispeech @ 3bdde20),
written function-by-function with the reference file:line each function follows.  It is pinned
against golden vectors produced by the real reference modules (tests/golden/make_golden.py,
fixtures under tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; nothing under optispeech_b200/ does.
"""
