"""Template-style helpers on top of infer(): classify / embed / score.

Callers of the hot path in the reference (`sutro/templates/classification.py:11-117`,
`embed.py:8-53`, `evals.py:12-74`); SURVEY.md §8(f).1 asks that they run against the
local engine.  Signatures (names, order, defaults) mirror the reference so existing call
sites keep working; the bodies are written for this backend: each helper builds a system
prompt and an `output_schema`, submits a detached job through `self.infer(...)` and
resolves it with `self.await_job_completion(...)`.  Prompts are this repo's wording.
`rank` / `elo` are not provided (upstream's pandas branches do not run).
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple, Union

import pandas as pd

from .common import EmbeddingModelOptions, ModelOptions, pl
from .interfaces import BaseSutroClient


class Templates(BaseSutroClient):
    def classify(
        self,
        data,
        classes: Union[Dict[str, str], List[str]],
        model: ModelOptions = "gemma-3-12b-it",
        job_priority: int = 0,
        name: Union[str, List[str]] = None,
        description: Union[str, List[str]] = None,
        output_column: str = "inference_result",
        column: Union[str, List[str]] = None,
        truncate_rows: bool = True,
        include_scratchpad: bool = False,
    ):
        """Assign every row to one of `classes` (list of labels or {label: description}).
        The schema gives the model a short scratchpad and then constrains the answer to the
        label set, so every row's `classification` is one of the labels by construction."""
        if not classes:
            raise ValueError("classes must name at least one label")
        labels = list(classes.keys()) if isinstance(classes, dict) else list(classes)
        lines = [f"* {k} — {v}" for k, v in classes.items()] if isinstance(classes, dict) \
            else [f"* {c}" for c in labels]
        system_prompt = ("Decide which single label best describes the input.\n"
                         "Labels:\n" + "\n".join(lines) + "\n"
                         "Think briefly in `scratchpad`, then give exactly one label in "
                         "`classification`.")
        schema = {"type": "object",
                  "properties": {"scratchpad": {"type": "string", "maxLength": 96},
                                 "classification": {"type": "string", "enum": labels}},
                  "required": ["scratchpad", "classification"]}
        job_id = self.infer(data, model, name, description, system_prompt=system_prompt,
                            output_schema=schema, column=column, output_column=output_column,
                            job_priority=job_priority, truncate_rows=truncate_rows,
                            stay_attached=False)
        results = self.await_job_completion(job_id, output_column=output_column)
        if results is None or include_scratchpad:
            return results
        return results[["classification"]].rename(columns={"classification": output_column})

    def embed(
        self,
        data,
        model: EmbeddingModelOptions = "qwen-3-embedding-0.6b",
        job_priority: int = 0,
        name: Union[str, List[str]] = None,
        description: Union[str, List[str]] = None,
        output_column: str = "inference_result",
        column: Union[str, List[str]] = None,
        truncate_rows: bool = True,
    ):
        """One L2-normalised embedding vector per row (list[float] in `output_column`);
        `get_job_embeddings(job_id)` returns the same data as one fp32 array."""
        job_id = self.infer(data, model, name, description, column, output_column, job_priority,
                            truncate_rows=truncate_rows, stay_attached=False)
        return self.await_job_completion(job_id, output_column=output_column)

    def score(
        self,
        data,
        model: ModelOptions = "gemma-3-12b-it",
        job_priority: int = 0,
        name: Union[str, List[str]] = None,
        description: Union[str, List[str]] = None,
        column: Union[str, List[str]] = None,
        criteria: Union[str, List[str]] = None,
        score_column_name: str = "score",
        range: Tuple[int, int] = (0, 10),
    ):
        """LLM-as-a-judge integer score in [range[0], range[1]] per row; returns the input
        frame with `score_column_name` appended (or the results frame for list input)."""
        if criteria is None:
            raise ValueError("criteria must be given")
        crit = [criteria] if isinstance(criteria, str) else list(criteria)
        lo, hi = int(range[0]), int(range[1])
        if hi < lo:
            raise ValueError("range must be (low, high) with low <= high")
        system_prompt = ("You grade the input against these criteria: " + "; ".join(crit) +
                         f". Answer with one integer from {lo} to {hi}.")
        schema: Dict[str, Any] = {"type": "object",
                                  "properties": {score_column_name: {"type": "integer",
                                                                     "minimum": lo, "maximum": hi}},
                                  "required": [score_column_name]}
        job_id = self.infer(data=data, model=model, name=name, description=description,
                            column=column, system_prompt=system_prompt, output_schema=schema,
                            job_priority=job_priority, stay_attached=False)
        res = self.await_job_completion(job_id)
        if res is None:
            return None
        if isinstance(data, pd.DataFrame):
            return data.assign(**{score_column_name: res[score_column_name].values})
        if pl is not None and isinstance(data, pl.DataFrame):
            return data.with_columns(pl.Series(score_column_name, list(res[score_column_name])))
        return res
