"""CPU oracle for the multimodal-organ-segmentation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package (``multimodal-organ-segmentation_b200/``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and there only as the checker or as the
reported CPU baseline.

The oracle is a plain fp32 (optionally fp64) restatement of the reference's
algorithm written against ``torch.nn.functional`` on the CPU: it takes a
``state_dict`` with the reference's parameter names and evaluates the same
arithmetic the reference's ``nn.Module`` tree evaluates.

Parity status
-------------
* UNet3D / DualEncoder / fusion modules / losses / DiceMetric: PINNED.  The
  restatement is checked (tests/test_oracle_vs_reference.py, run in the build
  container where /root/reference exists) against the reference's own modules
  imported read-only, and against golden vectors generated from the reference
  by ``tests/golden/make_golden.py`` (committed under tests/golden/).
* Sliding-window inference: PARITY UNPINNED.  The arithmetic lives in MONAI
  (``monai.inferers.sliding_window_inference``, requirement ``monai>=1.3.0`` —
  a floor, no lock file; MONAI is not vendored and not installed here, and the
  reference has no tests or golden vectors for it).  ``sliding_window.py``
  restates MONAI's published algorithm and is anchored
  only on the reference's call site (src/trainer/trainer.py:381-392) and on the
  algorithm's own invariants.
"""
