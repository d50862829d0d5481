"""Per-cycle block-size policy (reference: benchmark_dynamic_schedule.py:54-257).

The engine is built for the largest candidate block; a smaller block is just a shorter `blk_len` on the device
(rows past it are skipped by the K/V writer, masked out of the attention and ignored by the acceptance kernel), so
choosing the block size per cycle costs one 4-byte write. This module is the host policy only: it scores every
candidate block size by an exponentially weighted estimate of committed tokens per second, tau / cycle time, and
moves to a better candidate conservatively.

Policy (the reference's "ewma" mode, restated):
  * the first `warmup_cycles` cycles walk the candidates round-robin to seed the estimates;
  * every `probe_interval` cycles after that, one cycle is spent on a non-current candidate (round-robin) so stale
    estimates get refreshed;
  * after each cycle the estimates of the size that ran are updated (tail cycles shorter than every candidate are
    ignored); when the best-scoring size beats the current one by more than `switch_margin` (relative) for
    `required_streak` consecutive updates, it becomes current and switching is frozen for `cooldown_cycles`;
  * `low_accept_streak` consecutive cycles of the current size with tau / size below `low_accept_threshold` step
    down to the next smaller candidate at once.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence


class EwmaBlockScheduler:
    def __init__(self, candidates: Sequence[int], *, warmup_cycles: int = 0, ewma_alpha: float = 0.3,
                 switch_margin: float = 0.05, required_streak: int = 2, cooldown_cycles: int = 4,
                 probe_interval: int = 0, low_accept_threshold: float = 0.0, low_accept_streak: int = 3):
        cand = sorted({int(c) for c in candidates})
        if not cand or cand[0] < 2:
            raise ValueError("candidate block sizes must be >= 2")
        if not 0.0 < ewma_alpha <= 1.0:
            raise ValueError("ewma_alpha must be in (0, 1]")
        self.candidates: List[int] = cand
        self.current = cand[-1]
        self.alpha = float(ewma_alpha)
        self.warmup_cycles = max(0, int(warmup_cycles))
        self.switch_margin = max(0.0, float(switch_margin))
        self.required_streak = max(1, int(required_streak))
        self.cooldown_cycles = max(0, int(cooldown_cycles))
        self.probe_interval = max(0, int(probe_interval))
        self.low_accept_threshold = float(low_accept_threshold)
        self.low_accept_streak = max(1, int(low_accept_streak))
        self.tau: Dict[int, Optional[float]] = {b: None for b in cand}
        self.cycle_s: Dict[int, Optional[float]] = {b: None for b in cand}
        self.n_obs: Dict[int, int] = {b: 0 for b in cand}
        self._cooldown = 0
        self._challenger = self.current
        self._streak = 0
        self._low = 0
        self._probe_at = 0
        self.history: List[int] = []

    @property
    def max_block_size(self) -> int:
        return self.candidates[-1]

    def score(self, b: int) -> Optional[float]:
        """Estimated committed tokens per second at block size b."""
        if self.tau[b] is None:
            return None
        return self.tau[b] / max(1e-12, self.cycle_s[b])

    def _blend(self, old: Optional[float], new: float) -> float:
        return float(new) if old is None else (1.0 - self.alpha) * old + self.alpha * float(new)

    def select(self, cycle_idx: int) -> int:
        """Block size to run in cycle `cycle_idx`."""
        if cycle_idx < self.warmup_cycles:
            b = self.candidates[cycle_idx % len(self.candidates)]
        elif self.probe_interval and (cycle_idx - self.warmup_cycles) % self.probe_interval == 0:
            b = self.current
            for _ in range(len(self.candidates)):  # next candidate in the probe rotation that is not current
                c = self.candidates[self._probe_at % len(self.candidates)]
                self._probe_at += 1
                if c != self.current:
                    b = c
                    break
        else:
            b = self.current
        self.history.append(b)
        return b

    def update(self, *, tau: float, cycle_s: float, effective_bs: int, cycle_idx: int) -> None:
        b = int(effective_bs)
        if b not in self.tau:  # a clamped tail block: not evidence about any candidate
            return
        self.tau[b] = self._blend(self.tau[b], tau)
        self.cycle_s[b] = self._blend(self.cycle_s[b], cycle_s)
        self.n_obs[b] += 1

        if b == self.current and tau / max(1.0, float(b)) < self.low_accept_threshold:
            self._low += 1
        else:
            self._low = 0
        if self._low >= self.low_accept_streak:
            self._low = 0
            i = self.candidates.index(self.current)
            if i > 0:
                self.current = self.candidates[i - 1]
                self._challenger, self._streak, self._cooldown = self.current, 0, self.cooldown_cycles

        if cycle_idx < self.warmup_cycles:
            return
        if self._cooldown > 0:
            self._cooldown -= 1
            return
        best, best_score = None, None
        for c in self.candidates:  # ties go to the smaller block
            sc = self.score(c)
            if sc is not None and (best_score is None or sc > best_score):
                best, best_score = c, sc
        if best is None:
            return
        cur = self.score(self.current)
        # a current size that has never run has no score: nothing to compare against yet
        gain = -1.0 if cur is None else (best_score - cur) / max(1e-12, abs(cur))
        if best == self.current or gain <= self.switch_margin:
            self._challenger, self._streak = self.current, 0
            return
        self._streak = self._streak + 1 if best == self._challenger else 1
        self._challenger = best
        if self._streak >= self.required_streak:
            self.current = best
            self._streak, self._cooldown = 0, self.cooldown_cycles
