"""CPU oracle for the RestoraGen Stable-Diffusion sampling loop.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
reported CPU baseline -- never as the thing shipped.  The product package
(``image_restoration_and_enhancement_b200``) does not import ``oracle`` anywhere
and fails loudly when its CUDA library is missing.

What it restates
----------------
The reference (``/root/reference/src/inference.py``) delegates its whole hot
path to the third-party ``diffusers`` package (``requirements.txt:8``
``diffusers>=0.30``; checkpoints stamped ``_diffusers_version: 0.35.2`` in
``outputs/models/denoising/best/model_index.json:3``).  ``diffusers`` is neither
vendored under ``/root/reference`` nor installed in this image and there is no
network, so the real implementation cannot be imported or compiled.  Every
module here is a plain fp32 PyTorch restatement of the published diffusers
0.35.x algorithm, with state-dict key names identical to diffusers' so a real
SD-1.5 checkpoint can be loaded as a cross-check if one ever becomes available.

PARITY UNPINNED at the diffusers boundary: the reference ships no tests and no
golden vectors.  The pins that do exist and are checked in
``tests/test_oracle_*.py``:

* parameter counts: UNet 859,520,964 -- the figure the reference logged in
  ``outputs/models/colorization/training_colorization.log:30``; 9-ch inpaint
  UNet 859,535,364; VAE 83,653,863;
* scheduler tables: beta/alpha-bar endpoints from
  ``outputs/models/*/best/scheduler/scheduler_config.json``, the timestep lists
  of SURVEY.md section 8(d);
* tokenizer ids for the reference's default prompts (``src/inference.py:86-91``)
  via the shipped ``tokenizer/`` files (fixtures in ``tests/golden``).
"""
