"""Algorithm interface (master / RPC_ convention) and the built-in algorithms.

``IMAGES`` maps "image" names (what a task names) to python modules -- the process-runtime
equivalent of pulling an algorithm container.  Nodes may extend/override it with an
``algorithms: {image: module}`` block in their configuration."""
IMAGES = {
    # vantage6's canonical demo algorithm (column average)
    "harbor2.vantage6.ai/demo/average": "vantage6_b200.algorithm.builtin.average",
    "v6-average-py": "vantage6_b200.algorithm.builtin.average",
    "v6b200/average": "vantage6_b200.algorithm.builtin.average",
    # descriptive statistics on tabular node data (control plane only: a few numbers per column)
    "v6b200/summary": "vantage6_b200.algorithm.builtin.summary",
    "v6-summary-py": "vantage6_b200.algorithm.builtin.summary",
    "v6b200/crosstab": "vantage6_b200.algorithm.builtin.crosstab",
    "v6-crosstab-py": "vantage6_b200.algorithm.builtin.crosstab",
    "v6b200/kaplan-meier": "vantage6_b200.algorithm.builtin.kaplan_meier",
    "v6-kaplan-meier-py": "vantage6_b200.algorithm.builtin.kaplan_meier",
    "v6b200/correlation": "vantage6_b200.algorithm.builtin.correlation",
    "v6b200/coxph": "vantage6_b200.algorithm.builtin.coxph",
    "v6-coxph-py": "vantage6_b200.algorithm.builtin.coxph",
    # BASELINE config 1: weighted mean of a parameter vector
    "v6b200/weighted-mean": "vantage6_b200.algorithm.builtin.weighted_mean",
    # BASELINE configs 2-4: FedAvg over NVLink symmetric memory
    "v6b200/fedavg": "vantage6_b200.algorithm.builtin.fedavg",
    # BASELINE config 5: federated logistic-regression GLM
    "v6b200/glm": "vantage6_b200.algorithm.builtin.glm",
}


def resolve_image(image: str, extra: dict | None = None, allow_modules: bool = False) -> str:
    table = dict(IMAGES)
    table.update(extra or {})
    if image in table:
        return table[image]
    base = image.split(":")[0] if not image.startswith("module:") else image
    if base in table:
        return table[base]
    if image.startswith("module:") and allow_modules:
        return image[len("module:"):]
    raise KeyError(f"unknown algorithm image {image!r} (known: {sorted(table)})")
