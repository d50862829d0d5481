"""Golden vectors of the reference's multi-candidate builder and candidate choice
(benchmark_candidate_solutions.py:181-249 and :590-607).

Run in the build container only (needs /root/reference):  python tests/golden/make_cand_golden.py
`build_fixed_prefix_rank_candidates` is executed from the reference's source as is (only that function statement:
the module around it imports datasets / tokenizers). The choice rule is the reference's own three expression lines,
evaluated on the recorded inputs. Writes tests/golden/candidates.pt; tests/test_oracle_golden.py replays it.
"""
import ast
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/benchmark_candidate_solutions.py"


def main():
    tree = ast.parse(open(SRC).read())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "build_fixed_prefix_rank_candidates")
    ns = {"torch": torch}
    exec(compile(ast.Module([node], []), SRC, "exec"), ns)
    build = ns["build_fixed_prefix_rank_candidates"]
    g = torch.Generator().manual_seed(11)
    cases = []
    # (eff, V, prefix, rank_top_k, max_candidates, logits dtype, accepted lengths to force per candidate or None)
    for (eff, V, prefix, topk, maxc, dtype, force) in [
            (16, 500, 2, 4, 4, torch.bfloat16, None), (16, 500, 5, 4, 3, torch.float32, None),
            (8, 64, 0, 2, 4, torch.bfloat16, None), (5, 100, 9, 4, 4, torch.bfloat16, None),
            (16, 300, 2, 1, 4, torch.bfloat16, None), (3, 50, 1, 4, 4, torch.float32, None),
            (16, 500, 2, 4, 4, torch.bfloat16, [4, 6, 3, 13]), (16, 500, 2, 4, 4, torch.bfloat16, [5, 12, 12, 2]),
            (16, 500, 3, 4, 4, torch.bfloat16, [9, 9, 9, 9]), (12, 200, 2, 4, 2, torch.float32, [3, 11])]:
        logits = (torch.randn(1, eff - 1, V, generator=g) * 3).to(dtype)
        base = torch.cat([torch.tensor([[7]]), logits.argmax(-1)], dim=1)
        cands, meta, suffix = build(base, logits, prefix, topk, maxc)
        stacked = torch.cat(cands, dim=0)
        scores = [float(m["draft_score"]) for m in meta]
        # a posterior per candidate that agrees with a random-length prefix of it
        post = torch.randint(0, V, (stacked.shape[0], eff), generator=g)
        for k in range(stacked.shape[0]):
            n_ok = int(torch.randint(0, eff, (1,), generator=g)) if force is None else force[k]
            post[k, :n_ok] = stacked[k, 1:n_ok + 1]
            if n_ok < eff - 1:
                post[k, n_ok] = (stacked[k, n_ok + 1] + 1) % V  # and certainly not one more
        # the reference's choice (benchmark_candidate_solutions.py:590-607)
        acceptance_lengths_all = (stacked[:, 1:] == post[:, :-1]).cumprod(dim=1).sum(dim=1)
        tau_all = acceptance_lengths_all + 1
        draft_scores = torch.tensor(scores, dtype=torch.float32)
        candidate_indices = torch.arange(stacked.shape[0], dtype=torch.float32)
        composite = tau_all.float() * 1e6 + draft_scores - candidate_indices * 1e-3
        chosen = int(torch.argmax(composite).item())
        cases.append(dict(eff=eff, V=V, prefix=prefix, topk=topk, maxc=maxc, logits=logits, base=base, cands=stacked,
                          scores=scores, suffix=suffix, posterior=post, chosen=chosen,
                          acc=[int(x) for x in acceptance_lengths_all.tolist()]))
    torch.save(cases, os.path.join(HERE, "candidates.pt"))
    print("wrote", len(cases), "cases;", [(c["cands"].shape[0], c["chosen"]) for c in cases])


if __name__ == "__main__":
    main()
