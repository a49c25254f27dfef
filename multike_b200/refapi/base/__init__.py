"""Mirror of the parts of code/base/ that sit on the GPU path: evaluation.py, alignment.py
(greedy_alignment) and batch.py (generate_neighbours)."""
