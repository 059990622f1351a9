"""Oracle package: TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference hot path (tsurumeso/vocal-remover @ 99f92fe):
STFT -> CascadedNet forward -> mask -> inverse STFT, plus the sample-rate conversion in
front of it (``resample_oracle``).  Nothing under ``oracle/``
is part of the product path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import it, and
only as the checker / CPU baseline, never as the thing shipped.
"""
