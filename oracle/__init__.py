"""CPU oracle for the neurecon ray-marched SDF volume-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline.  The product package ``neurecon_b200``
never imports this package and fails loudly when its CUDA library is missing.

The oracle is a from-scratch restatement, on plain CPU torch tensors (fp32 by
default, fp64 on request), of the algorithms in the reference
(SuwoongHeo/neurecon, paths relative to /root/reference):

* ``oracle.nets``      -- models/base.py:14-81,118-129,243-282,372-391,426-453
* ``oracle.sampling``  -- utils/rend_util.py:167-234,255-327
* ``oracle.neus``      -- models/frameworks/neus.py:21-70,118-397
* ``oracle.volsdf``    -- models/frameworks/volsdf.py:16-272,306-331,334-551
* ``oracle.unisurf``   -- models/frameworks/unisurf.py:34-62,64-283 and
                          models/ray_casting.py:11-160

Parity pinning: the reference ships no tests, golden vectors or known-answer
fixtures for this path, so the oracle is pinned
against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (which imports the unmodified reference from
/root/reference through ``oracle.ref_loader``) and committed under
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks the oracle
against those vectors on every CPU run.
"""
