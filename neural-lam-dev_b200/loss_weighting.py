"""Per-variable loss weights (/root/reference/neural_lam/loss_weighting.py:
8-106): uniform 1/n, or a manual {variable: weight} table that must cover
exactly the datastore's state variables."""


def get_state_feature_weighting(config, datastore):
    names = datastore.get_vars_names(category="state")
    wcfg = config.training.state_feature_weighting
    manual = getattr(wcfg, "weights", None)
    if manual is None:
        return [1.0 / len(names)] * len(names)
    if set(manual) != set(names):
        raise ValueError(
            "State feature weights must be provided for each state feature in "
            f"the datastore ({names}); missing {set(names) - set(manual)}, "
            f"unknown {set(manual) - set(names)}"
        )
    return [manual[n] for n in names]
