"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the DiverGen generation hot path.

PARITY UNPINNED: the reference's arithmetic for this path lives in the third-party `diffusers`
package, which /root/reference neither vendors nor pins (absent from DiverGen/requirements.txt:1-17)
and which cannot be imported or installed here.  The reference holds no test, golden vector or
fixture for this path.  This package restates the published algorithm
(diffusers UNet2DConditionModel / DDIMScheduler / StableDiffusionPipeline as called from
DiverGen/generation/txt2img_diffusers_stages_from_txt.py:139-323) and is pinned only by structural
self-checks: parameter counts, state-dict key census, scheduler known answers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under divergen_b200/ may import it.
"""
