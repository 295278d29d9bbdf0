"""CPU oracle for the denoising-loop hot path of milesgray/complex_prompt_diffusion.

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package, and only as
the checker (or the timed CPU baseline), never as part of the product path.  The product package
``complex_prompt_diffusion_b200`` must never import ``oracle``.

It is a pure-torch (CPU, fp32/fp64) restatement of the reference arithmetic; every function cites
the reference ``file:line`` it follows (paths relative to ``/root/reference``).  The reference has no
tests or golden vectors, and cannot be imported as committed; parity is pinned
by running the *shimmed* reference in the build container (``oracle/ref_shim.py`` +
``oracle/make_golden.py``) and freezing its outputs as fixtures in ``tests/golden``, which
``tests/test_oracle_golden.py`` checks the oracle against.

Reference defects the restatement resolves:
  D1 SigmaScheduler has no append_zero (discrete.py:107)            -> implemented as in k.py:575-576.
  D2 SigmaScheduler.sigmas is None/overwritten (discrete.py:15-19,107) -> KScheduler's separate 1000-entry
     training table is used for sigma_to_t / t_to_sigma / linear (k.py:98,268-279).
  D3 sigma passed twice to _process_conditioning (denoiser.py:508 vs :530) -> passed once.
  D4 SpatialTransformer does not accept use_linear/use_checkpoint (unet.py:592-596) -> dropped for SD-1.x;
     SD-2.x linear proj_in/out restated from the yaml intent (v2-inference.yaml:34).
  D5 CUDA memory probing in CrossAttention.forward (attention.py:301-306) -> single slice (steps == 1).
  D6 hard-coded .cuda() -> device agnostic.
  D7 image batch > 1 unsupported (denoiser.py:390-391) -> B images = B independent batch-1 trajectories.
  D8 gamma > 0 branch uses sigma*2 (denoiser.py:536) -> gamma is kept 0 (the default).
"""
