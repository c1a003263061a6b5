"""CPU oracle for the ViMoCLIP per-frame encoding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``vimo-clip_b200/``) may import,
call, link or execute anything in this directory; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do,
and there only as the checker or as the reported CPU baseline.

What it is: a plain fp32 PyTorch / numpy restatement of the reference's algorithm for this
path, each function citing the reference ``file:line`` it follows:

* ``clip_shim``  -- the un-vendored OpenAI ``clip`` package (openai/CLIP @ dcba3cb2, pinned at
  ``requirements.txt:5``): ``VisionTransformer`` + ``_transform`` restated from its published
  architecture; lets the reference's own ``models/student_model*.py`` run unmodified.
* ``student``    -- ``FlowStudentModel`` / ``FrameDiffStudentModel`` forward.
* ``tfam``       -- ``AttentionLayer`` / ``AMO_CLIP``.
* ``prologue``   -- uint8 wrap / normalise / patchify and the OpenCV frame difference (numpy).
* ``indexing``   -- frame sampling, sparse sampling, padding masks (integer, bit-exact).
* ``losses``     -- cosine / MSE distillation loss, BCE classification loss.
* ``weights``    -- seeded weight generators shared by the golden-vector script and the tests.

Parity pinning: the reference ships no tests, golden vectors or fixtures, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container:
``oracle/make_golden.py`` imports the reference's Python files from ``/root/reference`` (and the
installed HF ``transformers`` CLIP tower / torchvision / OpenCV, which are the reference's
third-party numerics) and freezes seeded input/output vectors under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every oracle function against those fixtures.
"""
