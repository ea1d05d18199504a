"""TEST INFRASTRUCTURE ONLY — the parity oracle for uavsal-b200.

Nothing under ``iip_uavsal_saliency_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may use it,
and there only as the checker or the reported CPU baseline — never as the product path.

Contents
  shim.py        import the unmodified reference from /root/reference (authoring container only)
  cpu_ref.py     functional CPU restatement of the reference hot path (torch CPU fp32 primitives)
  synth.py       re-export of the seeded synthetic-input generators (iip_uavsal_saliency_b200/synth.py)
  make_golden.py regenerates tests/golden/*.npz by running the real reference through shim.py
  make_ckpt_fixture.py  checkpoint-loader fixtures (reference modules pickled in the authoring container)

Parity pin: the reference ships no tests or golden vectors, so cpu_ref.py is pinned
against outputs of the reference itself executed in the authoring container (make_golden.py →
tests/golden/), checked by tests/test_oracle_golden.py.
"""
