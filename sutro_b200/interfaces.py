"""Client interface and job-status enum of the local backend.

Kept signature-for-signature with the reference's `sutro/interfaces.py:11-91`
(north_star: "sutro.interfaces stays") so template mixins written against
`BaseSutroClient` work unchanged; only the frame types are duck-typed because polars
is an optional dependency here.
"""
from __future__ import annotations

from enum import Enum
from typing import Any, Dict, List, Optional, Type, Union

from .common import ModelOptions


class BaseSutroClient:
    """Declares what template mixins may call (reference: sutro/interfaces.py:11-66)."""

    def infer(
        self,
        data: Any,
        model: Union[ModelOptions, List[ModelOptions]] = "gemma-3-12b-it",
        name: Union[str, List[str]] = None,
        description: Union[str, List[str]] = None,
        column: Union[str, List[str]] = None,
        output_column: str = "inference_result",
        job_priority: int = 0,
        output_schema: Union[Dict[str, Any], Type[Any]] = None,
        sampling_params: dict = None,
        system_prompt: str = None,
        dry_run: bool = False,
        stay_attached: Optional[bool] = None,
        random_seed_per_input: bool = False,
        truncate_rows: bool = True,
    ) -> Any: ...

    def await_job_completion(
        self,
        job_id: str,
        timeout: Optional[int] = 7200,
        obtain_results: bool = True,
        output_column: str = "inference_result",
        is_cost_estimate: bool = False,
    ) -> Any: ...


class JobStatus(str, Enum):
    """Job states (reference: sutro/interfaces.py:69-91)."""

    UNKNOWN = "UNKNOWN"
    QUEUED = "QUEUED"
    STARTING = "STARTING"
    RUNNING = "RUNNING"
    SUCCEEDED = "SUCCEEDED"
    CANCELLING = "CANCELLING"
    CANCELLED = "CANCELLED"
    FAILED = "FAILED"

    @classmethod
    def terminal_statuses(cls) -> list["JobStatus"]:
        return [cls.SUCCEEDED, cls.FAILED, cls.CANCELLING, cls.CANCELLED]

    def is_terminal(self) -> bool:
        return self in self.terminal_statuses()
