"""CPU oracle for the learned-lifting / tree-entropy hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
timed CPU baseline), never as the thing shipped.  The product package
(``imagecompressionlearnedliftingandlearnedtreebasedmodels_b200``) never imports
this package and fails loudly when its CUDA library is missing.

What is here
------------
* ``thirdparty.py``  restatement of the arithmetic of the reference's
  un-vendored third-party dependencies that sit on the hot path
  (compressai==1.2.1: GaussianConditional / EntropyBottleneck / GDN /
  LowerBound / colour transforms; pytorch_wavelets (unpinned) + PyWavelets
  1.3.0 ``bior4.4`` periodised filter bank).  Published algorithms, anchored
  on the reference's call sites (file:line in every docstring).
* ``lifting.py``, ``dwt97.py``, ``subband_ae.py``, ``entropy.py``, ``model.py``
  a functional (state_dict in, tensors out) restatement of the reference's
  own modules on the path, each function citing the reference file:line.
* ``shims/``  package-shaped wrappers over ``thirdparty.py`` so that the
  UNMODIFIED reference (``/root/reference``) imports in the dev container.
  ``refload.py`` does that import; ``tests/golden/make_golden.py`` uses it to
  write the committed golden vectors.

Parity pin
----------
The reference ships no tests and no golden vectors, so
the pin is: outputs of the reference's own modules run in the dev container on
top of ``shims/`` (the only way they can run offline), committed under
``tests/golden/`` with the generating script.  The restatement in this package
is checked against those fixtures by ``tests/test_oracle_golden.py``.  The
third-party arithmetic itself (compressai / pytorch_wavelets) cannot be run
here, so for those pieces parity is pinned to the published algorithm plus the
in-repo numeric anchors (9/7 taps ``lifting_dwt_nets.py:415-418``, lifting
coefficients ``:431-432``, ``scale_bound=0.11``) and to scipy cross-checks:
"third-party parity unpinned by upstream fixtures".
"""
