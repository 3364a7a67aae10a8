"""CPU oracle for the MambaTTSDecoder hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``mamba_tts_project_b200/`` may import this package.  The only
legal importers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only as
the checker (or as the timed CPU baseline), never as the shipped compute path.

What it restates
----------------
* ``ssm_ref``      -- the arithmetic of the third-party ``mamba_ssm`` /
  ``causal_conv1d`` packages that ``/root/reference/mamba_decoder.py:4,29,61,63``
  reaches through ``Mamba(d_model)``: ``selective_scan_ref``,
  ``selective_state_update_ref``, ``causal_conv1d_ref``,
  ``causal_conv1d_update_ref`` (published algorithm of state-spaces/mamba and
  Dao-AILab/causal-conv1d; NOT vendored and NOT version-pinned by the reference:
  ``environment.yml:1-150`` has no entry for either).
* ``mamba_ref``    -- the non-fused ``Mamba.forward`` / ``Mamba.step`` graph with
  the ``(out, state)`` contract the reference decoder expects
  (``mamba_decoder.py:9-15,61,63``).
* ``decoder_ref``  -- behaviour of ``mamba_decoder.py:25-256`` with the block above
  injected.

Parity pin
----------
The reference holds no test, golden vector or fixture for this path
(SURVEY.md section 4 / 8c): **parity is unpinned by the reference itself**.  The
strongest pin available offline is an independent implementation of the same
published algorithm that ships in this image: HuggingFace ``transformers``
``MambaMixer.slow_forward`` (``models/mamba/modeling_mamba.py``).
``oracle/make_golden.py`` imports it, runs it on seeded inputs and commits the
vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
oracle against them on every run.
"""
