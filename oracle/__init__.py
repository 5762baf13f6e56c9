"""CPU oracle for the VisTracker hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; ``vistracker_b200`` never does (tests/test_boundary.py greps for it).

Each function restates one piece of the reference algorithm in plain PyTorch-CPU / numpy and cites the
reference ``file:line`` it follows.  Parity status:

* ``sifnet_ref``  -- PINNED: checked against outputs of the unmodified reference classes
  (``model.CHORETriplaneVisibility``) imported from /root/reference in the build container; vectors in
  ``tests/golden/sifnet_*.npz``, generator ``tests/golden/make_golden.py``.
* ``smpl_ref``    -- PINNED the same way against the unmodified ``SMPL_Layer.forward`` on a synthetic
  SMPL-H-shaped model (the real ``SMPLH_male.pkl`` is not redistributable).
"""
