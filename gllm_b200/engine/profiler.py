"""torch.profiler control, started/stopped on every rank by a control command
(reference: gllm/profiler_mixin.py:12-98). Traces: gzip chrome trace per rank under
`$GLLM_TORCH_PROFILER_DIR/trace_session_<ts>/trace_rank<r>_<ts>.json.gz`. NVTX ranges are pushed
around every engine step while profiling so Nsight timelines show batch boundaries."""
from __future__ import annotations

import gzip
import os
import shutil
import time

import torch

from gllm_b200.utils.logging import logger


class ProfilerMixin:
    def init_profiler(self):
        self._prof = None
        self._prof_dir = None
        self.profiler_root = os.environ.get("GLLM_TORCH_PROFILER_DIR", "/tmp")

    @property
    def profiling(self) -> bool:
        return getattr(self, "_prof", None) is not None

    def start_profile(self, session_dir=None):
        if self._prof is not None:
            return
        ts = time.strftime("%Y%m%d_%H%M%S")
        self._prof_dir = session_dir or os.path.join(self.profiler_root, f"trace_session_{ts}")
        os.makedirs(self._prof_dir, exist_ok=True)
        acts = [torch.profiler.ProfilerActivity.CPU]
        if torch.cuda.is_available():
            acts.append(torch.profiler.ProfilerActivity.CUDA)
        self._prof = torch.profiler.profile(activities=acts, record_shapes=True, with_stack=True,
                                            profile_memory=False)
        self._prof.start()
        logger.info("profiler started -> %s", self._prof_dir)

    def stop_profile(self):
        if self._prof is None:
            return None
        self._prof.stop()
        ts = time.strftime("%Y%m%d_%H%M%S")
        rank = getattr(self, "rank", 0)
        raw = os.path.join(self._prof_dir, f"trace_rank{rank}_{ts}.json")
        self._prof.export_chrome_trace(raw)
        with open(raw, "rb") as fi, gzip.open(raw + ".gz", "wb") as fo:
            shutil.copyfileobj(fi, fo)
        os.remove(raw)
        self._prof = None
        logger.info("profiler trace written: %s.gz", raw)
        return raw + ".gz"
