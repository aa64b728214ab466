"""CPU oracle for the U-ResNet hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package.  The product
path (``u-resnet_b200/``) never imports it and has no CPU fallback.

PARITY -- what it is pinned to.  The reference (DeepLearnPhysics/u-resnet) ships
no golden vectors, no tests and no weights, and its arithmetic lives in
TensorFlow 1.x (``tf.contrib.slim``), which is neither vendored nor installable
here.  The oracle restates ``lib/uresnet.py:22-123``,
``lib/resnet_module.py:10-87`` and ``lib/ssnet.py:29-79`` with the slim/TF-1.x
op semantics written down in SURVEY.md section 3.4.  It is pinned (i) to the
reference's own Python, imported from /root/reference and EXECUTED in the build
container over an emulation of the TensorFlow calls it makes -- golden vectors
of those runs and their generating scripts are committed under
``tests/golden/`` (next paragraph); (ii) to the GraphDef dump TensorFlow itself
wrote for the reference (the paragraph after); (iii) to an independent naive
NumPy implementation (``oracle/naive.py``) and fp64 finite differences.
UNPINNED remains exactly one thing: the numerics of TensorFlow's own kernels,
which never ran here.

What the reference's OWN CODE pins: ``oracle/tf1_shim.py`` emulates the ~25
TensorFlow-1.x / slim calls that ``lib/uresnet.py``, ``lib/resnet_module.py``
and ``lib/ssnet.py`` make, so those three files are imported UNMODIFIED and
driven through their public API (construct / zero_gradients / accum_gradients
/ apply_gradients / inference) with torch-CPU evaluating the ops.  The graph
composition, variable creation order, loss reduction, accuracy expressions and
the accumulate / Adam protocol that run are the reference's, not a reading of
them; ``tests/golden/gref*.npz`` are outputs of such runs
(``make_reference_run_golden.py``) and ``tests/test_reference_run.py`` shows
the oracle reproducing them to fp64 round-off (loss 2e-16, float32-stored
logits and gradients bit-identical).  The semantics of the TensorFlow ops
themselves stay a restatement (the shim states them independently of the
oracle, e.g. conv-transpose literally as Conv2DBackpropInput), hence:
numerics of TensorFlow's kernels unpinned.

What else is pinned by the reference: the one artefact it produced itself, the
TensorBoard GraphDef dump under ``libs/uresnet_graph/`` (an older revision of
the model).  ``tests/golden/reference_graph.json`` is its committed extract and
``tests/test_reference_graph.py`` holds the oracle to it: the 58 filter names,
shapes and TensorFlow's own Xavier limits, the wiring / strides / padding of
every conv-type op, the skip map and concat operand order, the residual adds,
slim's lowering of conv2d_transpose to Conv2DBackpropInput, and the dumped
batch_norm and accuracy expressions EXECUTED against ``batch_norm`` /
``loss_and_metrics``.  Structure and op expansions are pinned that way; tensor
VALUES computed by TensorFlow exist nowhere.
"""
