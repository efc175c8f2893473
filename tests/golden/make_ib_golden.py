"""Golden vectors for the training-time in-batch-negative scoring, recorded by EXECUTING the reference.

`FLMRModelForRetrieval.compute_ib_loss_new` (src/models/flmr/models/flmr/modeling_flmr.py:1089-1125) and
`colbert_score_reduce` (src/models/flmr/models/flmr/flmr_utils.py:22-30) are pulled out of the unmodified reference
sources by their AST nodes and executed as they are (the surrounding module needs packages this image does not have);
`self.loss_fn` is the reference's `torch.nn.CrossEntropyLoss()` (modeling_flmr.py:720), wrapped only to capture the
score matrix and labels it is handed.  Run in the authoring container:  python tests/golden/make_ib_golden.py
"""
import ast
import os
import types

import numpy as np
import torch

REF = os.environ.get("PLAID_REFERENCE_ROOT", "/root/reference")
FLMR = os.path.join(REF, "src", "models", "flmr", "models", "flmr")
HERE = os.path.dirname(os.path.abspath(__file__))


def extract(path, name):
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise KeyError(name)


def main():
    ns = {"torch": torch}
    exec(extract(os.path.join(FLMR, "flmr_utils.py"), "colbert_score_reduce"), ns)
    import textwrap
    exec(textwrap.dedent(extract(os.path.join(FLMR, "modeling_flmr.py"), "compute_ib_loss_new")), ns)
    g = torch.Generator().manual_seed(20260)
    B, n_docs, Lq, Ld = 4, 3, 40, 23
    Q = torch.nn.functional.normalize(torch.randn(B, Lq, 128, generator=g), dim=-1).requires_grad_(True)
    D = torch.nn.functional.normalize(torch.randn(B * n_docs, Ld, 128, generator=g), dim=-1)
    lens = torch.randint(5, Ld + 1, (B * n_docs,), generator=g)
    lens[1] = Ld
    mask = (torch.arange(Ld).unsqueeze(0) < lens.unsqueeze(1)).unsqueeze(-1).float()    # [n, Ld, 1] like FLMR's context_mask
    D = (D * mask).requires_grad_(True)
    captured = {}
    ce = torch.nn.CrossEntropyLoss()

    def loss_fn(scores, labels):
        captured["scores"], captured["labels"] = scores.detach().clone(), labels.detach().clone()
        return ce(scores, labels)

    me = types.SimpleNamespace(loss_fn=loss_fn)
    loss = ns["compute_ib_loss_new"](me, Q, D, mask)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "ib_loss.npz"), Q=Q.detach().numpy(), D=D.detach().numpy(), mask=mask.numpy(),
                        scores=captured["scores"].numpy(), labels=captured["labels"].numpy(), loss=loss.detach().numpy(),
                        dQ=Q.grad.numpy(), dD=D.grad.numpy(), n_docs=n_docs)
    print("wrote ib_loss.npz: loss", float(loss), "scores", tuple(captured["scores"].shape))


if __name__ == "__main__":
    main()
