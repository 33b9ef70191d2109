"""Stand-in for pyrepseq==1.5 (only pyrepseq.nn.symdel, collapse.py:66,735)."""
