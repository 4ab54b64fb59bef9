"""GPU visual effects: the reference's output-stage effects (watsor/output/{copy,blend,draw}.py) as one CUDA pass."""
