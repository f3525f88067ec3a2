"""Fit -> save -> fit BPE -> save -> evaluate: the flow of the reference's train/train_beast.py:27-117 and
train/eval.py:11-114, runnable without the lerobot / hydra data plumbing.

    python -m beast_tokenizer_b200.train_beast --device cuda [--num-basis 10 --degree 4 --vocab-size 256 ...]

Same flags as the reference CLI; the dataloaders are synthetic (`--actions-len`, `--actions-dof`,
`--train-batches`, `--eval-batches`) because the reference's datasets live on private storage.
Outputs: the two checkpoint directories (byte-compatible with the reference's) and
<eval-results-dir>/<dataset>/{errors.json, stats.txt} + total_stats.json (no plots).
"""
import argparse
import json
from pathlib import Path
from typing import Any, Iterable, Iterator

import numpy as np

from .beast_bspline_bpe_tokenizer import BEASTBsplineBPETokenizer
from .beast_bspline_tokenizer import BEASTBsplineTokenizer
from .synth import SyntheticLoader


def _limit_batches(loader: Iterable[Any], max_batches) -> Iterator[Any]:
    """Reference train/train_beast.py:16-24: None or <= 0 means the whole loader."""
    if max_batches is None or max_batches <= 0:
        yield from loader
        return
    for i, batch in enumerate(loader):
        yield batch
        if i + 1 >= max_batches:
            break


def evaluate_from_path(dataloader, dataset_name: str, tokenizer_path: str, is_bpe_tokenizer: bool = True,
                       save_path: str = "eval_results", max_eval_samples: int = 12_500, device=None):
    save_dir = Path(save_path) / dataset_name
    save_dir.mkdir(parents=True, exist_ok=True)
    cls = BEASTBsplineBPETokenizer if is_bpe_tokenizer else BEASTBsplineTokenizer
    tokenizer = cls.from_pretrained(tokenizer_path, device=device)
    errors_l2, errors_l1, tokens_length = [], [], []
    for batch in dataloader:
        if len(errors_l2) >= max_eval_samples:
            break
        error_l2, error_l1, tokens = tokenizer.compute_reconstruction_error(batch["actions"], return_tokens=True)
        errors_l2.append(error_l2.item())
        errors_l1.append(error_l1.item())
        tokens_length.extend(len(row) for row in tokens)
    with open(save_dir / "errors.json", "w") as f:
        json.dump({"errors_l2": errors_l2, "errors_l1": errors_l1, "mean_tokens_length": tokens_length}, f)
    stats = {f"{fn.__name__}_{k}": float(fn(v)) for k, v in (("l2", errors_l2), ("l1", errors_l1))
             for fn in (np.mean, np.std, np.max, np.min)}
    stats = {k.replace("amax", "max").replace("amin", "min"): v for k, v in stats.items()}
    with open(save_dir / "stats.txt", "w") as f:
        for name, fn in (("Mean", np.mean), ("Std", np.std), ("Max", np.max), ("Min", np.min)):
            print(f"{name} tokens length:", fn(tokens_length), file=f)
        print("", file=f)
        for k in ("l2", "l1"):
            for name in ("mean", "std", "max", "min"):
                print(f"{name.capitalize()} reconstruction error {k}:", stats[f"{name}_{k}"], file=f)
            print("", file=f)
    return stats


def main(argv=None) -> None:
    parser = argparse.ArgumentParser(description="Train the base BEAST tokenizer and optionally the BEAST+BPE extension.")
    parser.add_argument("--batch-size", type=int, default=32)
    parser.add_argument("--num-basis", type=int, default=50)
    parser.add_argument("--vocab-size", type=int, default=1000)
    parser.add_argument("--degree", type=int, default=0)
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--fit-beast-max-samples", type=int, default=5_000)
    parser.add_argument("--fit-bpe-max-samples", type=int, default=25_000)
    parser.add_argument("--bpe-vocab-size", type=int, default=2048)
    parser.add_argument("--beast-checkpoint-dir", type=str, default="beast_tokenizer_checkpoint")
    parser.add_argument("--bpe-checkpoint-dir", type=str, default="beast_bpe_tokenizer_checkpoint")
    parser.add_argument("--eval-results-dir", type=str, default="eval_results")
    parser.add_argument("--max-eval-samples", type=int, default=12_500)
    group = parser.add_mutually_exclusive_group()
    group.add_argument("--train-bpe", dest="train_bpe", action="store_true")
    group.add_argument("--no-train-bpe", dest="train_bpe", action="store_false")
    parser.set_defaults(train_bpe=True)
    # synthetic stand-in for train/data.py: prepare_dataloaders
    parser.add_argument("--actions-len", type=int, default=10)
    parser.add_argument("--actions-dof", type=int, default=32)
    parser.add_argument("--train-batches", type=int, default=2_000)
    parser.add_argument("--eval-batches", type=int, default=200)
    parser.add_argument("--seed", type=int, default=0)
    args = parser.parse_args(argv)

    T, D = args.actions_len, args.actions_dof
    train = SyntheticLoader(args.train_batches, args.batch_size, T, D, seed0=args.seed)
    evals = {"synthetic": SyntheticLoader(args.eval_batches, args.batch_size, T, D, seed0=args.seed + 10_000_000)}

    tokenizer = BEASTBsplineTokenizer(num_basis=args.num_basis, vocab_size=args.vocab_size, degree_p=args.degree,
                                      num_dof=D, seq_len=T, init_pos=False, device=args.device)
    tokenizer.fit_parameters(train, max_samples=args.fit_beast_max_samples, verbose=False)
    Path(args.beast_checkpoint_dir).mkdir(parents=True, exist_ok=True)
    tokenizer.save_pretrained(args.beast_checkpoint_dir)
    print(f"Saved BEAST tokenizer to {args.beast_checkpoint_dir}")

    if not args.train_bpe:
        print("Skipping BPE training (use --train-bpe to enable).")
    else:
        bpe_tokenizer = BEASTBsplineBPETokenizer.from_beast(tokenizer, bpe_vocab_size=args.bpe_vocab_size)
        bpe_tokenizer.fit_from_trajectories(_limit_batches(train, args.fit_bpe_max_samples),
                                            max_sequences=args.fit_bpe_max_samples, show_progress=False)
        Path(args.bpe_checkpoint_dir).mkdir(parents=True, exist_ok=True)
        bpe_tokenizer.save_pretrained(args.bpe_checkpoint_dir)
        print(f"Saved BEAST+BPE tokenizer to {args.bpe_checkpoint_dir}")

    total_stats = {}
    for name, loader in evals.items():
        path = args.bpe_checkpoint_dir if args.train_bpe else args.beast_checkpoint_dir
        total_stats[name] = evaluate_from_path(loader, name, path, args.train_bpe, save_path=args.eval_results_dir,
                                               max_eval_samples=args.max_eval_samples, device=args.device)
    Path(args.eval_results_dir).mkdir(parents=True, exist_ok=True)
    with open(Path(args.eval_results_dir) / "total_stats.json", "w") as f:
        json.dump(total_stats, f, indent=4)
    print(json.dumps(total_stats))


if __name__ == "__main__":
    main()
