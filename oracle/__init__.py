"""CPU oracle for the DNNCancerAnnotator conv-stack hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU comparator), never as the thing shipped.  The product package
``dnncancerannotator_b200`` never imports this package and fails loudly when its
CUDA library is missing.

PARITY UNPINNED.  The reference (``/root/reference``) is pure Python on
TensorFlow 2.6 / Keras (``requirements.txt:2``); TensorFlow is not installable
in this environment (no wheel, no network, Python 3.12), the reference ships no
golden vectors, fixtures or tests for this path (its only test module,
``annotator/tests/test_region_metrics.py``, covers region metrics), and it has
no native sources that could be compiled into ``oracle/_ref``.  The oracle is
therefore an *independent restatement* of the reference's algorithm:

* ``ref_ops.py``     torch-CPU (oneDNN) fp32/fp64 functional restatement of each
                     Keras call site, with TF weight layouts (HWIO, transposed
                     conv ``[kh,kw,Cout,Cin]``) and TF/Keras default semantics.
* ``ref_models.py``  the model graphs of ``components.py`` / ``unet.py`` /
                     ``multiresunet.py`` and the loss of ``losses.py`` built
                     from ``ref_ops`` (autograd supplies the gradients, like
                     ``tf.GradientTape`` does in ``Model.fit``).
* ``ref_numpy.py``   a second, loop-level numpy-fp64 restatement (forward and
                     hand-derived backward) for tiny shapes, so that the two
                     restatements check each other, plus first-principles
                     known-answer tests in ``tests/test_oracle_*.py``.

Every function cites the reference ``file:line`` it follows; TensorFlow/Keras
defaults that are not visible in the reference source are marked
``[TF-semantics]``.
"""
