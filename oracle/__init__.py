"""CPU oracle for the DISSC inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import it, and there only as the checker / the CPU baseline being
reported -- never as the thing shipped.  ``dissc_b200`` never imports it.

Pinning status
--------------
* Generator / predictors / ``len_carryover_correction``: pinned against
  outputs of the reference itself (``/root/reference`` imported in the build
  container by ``tests/golden/make_golden.py``; vectors committed under
  ``tests/golden/``).  The reference ships no tests or golden vectors of its
  own, so this is the strongest pin available.
* HuBERT + k-means: **parity unpinned** -- the arithmetic lives in textlesslib
  (unpinned HEAD) and fairseq@dd106d95, neither of which is under
  ``/root/reference`` nor installable offline.  The restatement follows the
  published architecture and is cross-checked against torchaudio's
  ``hubert_base`` (random weights).
"""
