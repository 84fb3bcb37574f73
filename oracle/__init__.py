"""TEST INFRASTRUCTURE ONLY — CPU oracle for the LASS/AudioSep separation hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker or the timed CPU baseline — never as the thing shipped.

Contents
--------
``oracle/torchlibrosa/stft.py``  restatement of the un-vendored third-party dependency
                                 ``torchlibrosa==0.1.0`` (reference ``environment.yml:306``)
                                 that holds the STFT / ISTFT / magphase arithmetic
                                 (call sites: reference ``models/resunet.py:6,284-302,473,510``,
                                 ``models/base.py:6,80,84,146-149``).
``oracle/resunet_oracle.py``     functional fp32 restatement of ``ResUNet30.forward``
                                 (reference ``models/resunet.py:522-653``); travels to the GPU box.
``oracle/reference_loader.py``   imports the UNMODIFIED reference from ``/root/reference`` (only
                                 present in the build container) with the torchlibrosa restatement
                                 first on ``sys.path``; used to pin the restatement and to generate
                                 ``tests/golden``.
``oracle/factory.py``            seeded weights / inputs shared by every parity test.
``oracle/make_golden.py``        generator of ``tests/golden/*.npz`` (run in the build container).

Parity pin status: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself executed here (fixtures under
``tests/golden`` + ``tests/test_oracle_vs_reference.py``) and against fp64 ``torch.stft``.
"""
