"""Client protocol and job-state enum of the local backend.

The reference keeps these in `sutro/interfaces.py` (`BaseSutroClient` :11-66, `JobStatus`
:69-91) and north_star says that file "stays": template mixins are typed against it.  This
module states the same contract for the local package — identical method names, parameter
names, order and defaults, identical state names/values — without depending on polars.
`tests/test_sdk_plumbing.py` asserts the signatures against the reference's.
"""
from __future__ import annotations

import enum
from typing import Any, Optional

from .common import ModelOptions

DEFAULT_MODEL: ModelOptions = "gemma-3-12b-it"
DEFAULT_OUTPUT_COLUMN = "inference_result"


class BaseSutroClient:
    """What helpers layered on the client may rely on: submit a job, resolve a job."""

    def infer(self, data: Any, model: Any = DEFAULT_MODEL, name: Any = None, description: Any = None,
              column: Any = None, output_column: str = DEFAULT_OUTPUT_COLUMN, job_priority: int = 0,
              output_schema: Any = None, sampling_params: Optional[dict] = None,
              system_prompt: Optional[str] = None, dry_run: bool = False,
              stay_attached: Optional[bool] = None, random_seed_per_input: bool = False,
              truncate_rows: bool = True) -> Any:
        """Submit `data` (list, frame + `column`, or csv/parquet/txt path) for inference and
        return the job id; see `sutro_b200.sdk.Sutro.infer`."""
        raise NotImplementedError

    def await_job_completion(self, job_id: str, timeout: Optional[int] = 7200,
                             obtain_results: bool = True,
                             output_column: str = DEFAULT_OUTPUT_COLUMN,
                             is_cost_estimate: bool = False) -> Any:
        """Block until `job_id` is terminal; return its results frame (or None)."""
        raise NotImplementedError


_STATES = ("UNKNOWN", "QUEUED", "STARTING", "RUNNING", "SUCCEEDED", "CANCELLING", "CANCELLED",
           "FAILED")
_TERMINAL = ("SUCCEEDED", "FAILED", "CANCELLING", "CANCELLED")


class JobStatus(str, enum.Enum):
    """Lifecycle states a job reports (same names and string values as the reference)."""

    _ignore_ = ["_s"]
    _s = None
    for _s in _STATES:
        vars()[_s] = _s

    @classmethod
    def terminal_statuses(cls) -> list:
        return [cls[s] for s in _TERMINAL]

    def is_terminal(self) -> bool:
        return self.value in _TERMINAL
