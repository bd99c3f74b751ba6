"""Template-style helpers on top of infer(): classify / embed / score / rank / elo.

Callers of the hot path in the reference (`sutro/templates/classification.py:11-117`,
`embed.py:8-53`, `evals.py:12-74` score, `:77-180` rank, `:182-334` elo); SURVEY.md §8(f).1
asks that they run against the local engine.  Signatures (names, order, defaults) mirror the reference so existing call
sites keep working; the bodies are written for this backend: each helper builds a system
prompt and an `output_schema`, submits a detached job through `self.infer(...)` and
resolves it with `self.await_job_completion(...)`.  Prompts are this repo's wording.
`rank` constrains the answer to a permutation of the option labels (the reference only asks
for an array of strings); `elo` is the same Bradley-Terry fit, pinned against the reference's
own function by `tests/golden/plumbing.json`.  Pandas input works here (upstream's pandas
branches do not run).
"""
from __future__ import annotations

import itertools
import json
import math
from typing import Any, Dict, List, Tuple, Union

import numpy as np
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

    def rank(
        self,
        model: ModelOptions = "gemma-3-12b-it",
        job_priority: int = 0,
        name: Union[str, List[str]] = None,
        description: Union[str, List[str]] = None,
        data=None,
        option_labels: List[str] = None,
        criteria: Union[str, List[str]] = None,
        ranking_column_name: str = "ranking",
        run_elo: bool = True,
    ):
        """LLM-as-a-judge ranking of several options per row.  `data` is a list of rows (each a
        list with one text per label) or a frame whose columns include `option_labels`.
        Returns the frame (pandas in → pandas out) with `ranking_column_name` holding the
        labels ordered best → worst; with `run_elo` the Elo table over all rows is printed."""
        if data is None or not option_labels:
            raise ValueError("data and option_labels are required")
        if criteria is None:
            raise ValueError("criteria must be given")
        labels = [str(x) for x in option_labels]
        if len(set(labels)) != len(labels):
            raise ValueError("option_labels must be distinct")
        crit = [criteria] if isinstance(criteria, str) else list(criteria)
        if isinstance(data, list):
            if any(len(r) != len(labels) for r in data):
                raise ValueError("every row must hold one text per option label")
            frame = pd.DataFrame(data, columns=labels)
        elif pl is not None and isinstance(data, pl.DataFrame):
            frame = data.to_pandas()
        elif isinstance(data, pd.DataFrame):
            frame = data
        else:
            raise ValueError("data must be a list of lists or a DataFrame")
        missing = [c for c in labels if c not in frame.columns]
        if missing:
            raise ValueError(f"option_labels not found among the columns: {missing}")
        # "A: <text> B: <text> ..." — labels double as the column names, so the generic
        # column concatenation (which treats unknown names as separators) is not used
        parts = [frame[c].astype("string").fillna("") for c in labels]
        rows = [" ".join(f"{lab}: {col.iloc[i]}" for lab, col in zip(labels, parts))
                for i in range(len(frame))]
        system_prompt = ("You compare the labelled options in the input by these criteria: " +
                         "; ".join(crit) + ". The labels are " + ", ".join(labels) +
                         ". List every label once, best first.")
        if len(labels) <= 6:     # every permutation spelled out: the answer is a ranking by construction
            answer: Dict[str, Any] = {"enum": [list(p) for p in itertools.permutations(labels)]}
        else:
            answer = {"type": "array", "items": {"type": "string", "enum": labels},
                      "minItems": len(labels), "maxItems": len(labels)}
        schema = {"type": "object", "properties": {ranking_column_name: answer},
                  "required": [ranking_column_name]}
        job_id = self.infer(data=rows, model=model, name=name, description=description,
                            system_prompt=system_prompt, output_schema=schema,
                            job_priority=job_priority, stay_attached=False)
        res = self.await_job_completion(job_id, output_column=ranking_column_name)
        if res is None:
            return None
        rankings = [json.loads(x) if isinstance(x, str) else list(x)
                    for x in res[ranking_column_name]]
        if run_elo:
            table = self.elo(rankings)
            print(table[["elo", "wins", "losses", "matches"]].to_string())
        out = frame.assign(**{ranking_column_name: pd.Series(rankings, index=frame.index)})
        if pl is not None and isinstance(data, pl.DataFrame):
            return pl.from_pandas(out)
        return out

    @staticmethod
    def elo(
        data=None,
        column: Union[str, List[str]] = None,
        laplace: float = 0.5,
        max_iter: int = 1000,
        tol: float = 1e-8,
        elo_mean: float = 1500.0,
    ):
        """Elo-scaled Bradley-Terry abilities from ranked ballots (`rank` output).
        A ballot is a sequence best → worst; an element that is itself a list/tuple/set is a
        tie group.  Earlier groups beat later ones once per (member, member) pair, members of
        a group tie (half a win each way); every ordered pair then gets `laplace` pseudo-wins.
        Abilities are fitted by Hunter's MM iteration normalised to geometric mean 1, and
        reported as `elo = 400·log10(ability)` re-centred on `elo_mean`.  Returns a DataFrame
        indexed by label (columns ability, beta, elo, wins, losses, matches), best first."""
        if isinstance(data, pd.DataFrame) or (pl is not None and isinstance(data, pl.DataFrame)):
            if column is None:
                raise ValueError("column is required when ballots come in a DataFrame")
            ballots = list(data[column])
        else:
            ballots = list(data or [])
        wins: Dict[Tuple[str, str], float] = {}
        decisive = set()     # labels that appear in at least one won/lost pair

        def credit(a, b, amount):
            wins[(a, b)] = wins.get((a, b), 0.0) + amount

        for ballot in ballots:
            tiers = [[str(m) for m in t] if isinstance(t, (list, tuple, set)) else [str(t)]
                     for t in ballot if t is not None]
            for hi, better in enumerate(tiers):
                for worse in tiers[hi + 1:]:
                    for a in better:
                        for b in worse:
                            if a != b:
                                credit(a, b, 1.0)
                                decisive.update((a, b))
                for a, b in itertools.combinations(better, 2):
                    if a != b:   # a tie is half a win in both directions
                        credit(a, b, 0.5)
                        credit(b, a, 0.5)
        # the reference builds its label list from the decisive pairs only; labels that never
        # won or lost a comparison are left out of the fit
        order = sorted(decisive)
        pos = {n: i for i, n in enumerate(order)}
        m = len(order)
        w = np.zeros((m, m))
        for (a, b), c in wins.items():
            if a in pos and b in pos:
                w[pos[a], pos[b]] += c
        if laplace and laplace > 0:
            w += laplace
            np.fill_diagonal(w, 0.0)
        games = w + w.T
        played = games.sum(axis=1) > 0
        if m and not played.all():
            keep = np.flatnonzero(played)
            order = [order[i] for i in keep]
            w, games = w[np.ix_(keep, keep)], games[np.ix_(keep, keep)]
            m = len(order)
        ability = np.ones(m)
        won = w.sum(axis=1)
        for _ in range(max_iter if m else 0):
            prev = ability
            expect = (games / (prev[:, None] + prev[None, :] + 1e-12)).sum(axis=1)
            ability = np.where(expect > 0, won / np.where(expect > 0, expect, 1.0), prev)
            ability = ability / np.prod(ability) ** (1.0 / m)
            if np.max(np.abs(np.log(ability + 1e-12) - np.log(prev + 1e-12))) < tol:
                break
        beta = np.log(ability + 1e-12)
        elo = beta * (400.0 / math.log(10.0))
        elo = elo - (elo.mean() if m else 0.0) + elo_mean
        table = pd.DataFrame({"ability": ability, "beta": beta, "elo": elo, "wins": won,
                              "losses": w.sum(axis=0), "matches": games.sum(axis=1)}, index=order)
        return table.sort_values("elo", ascending=False)
