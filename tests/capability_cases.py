"""Classes that probe `AbstractModel.__init__`'s cross-checks of the capability methods
(/root/reference/embiggen/utils/abstract_models/abstract_model.py:32-133), written once and run
against BOTH base classes: the restatement in embiggen_b200/embedding_api.py (tests/test_embedder_api.py)
and the reference's own (tests/real_embiggen_probe.py).  The two must agree case by case.  The
classes live in a file because "implemented" is decided by `inspect.getsource`."""


def build_cases(AbstractModel):
    """name -> (callable that constructs / probes, expected outcome).  An outcome is the name of
    the exception type raised, or the repr of the value returned."""

    class Names:
        @classmethod
        def model_name(cls):
            return "probe"

        @classmethod
        def library_name(cls):
            return "probe library"

        @classmethod
        def task_name(cls):
            return "probe task"

    class UsesNothing(Names, AbstractModel):  # the minimal valid stochastic model
        def __init__(self, random_state=42):
            super().__init__(random_state)

        @classmethod
        def is_stocastic(cls):
            return True

        @classmethod
        def can_use_edge_weights(cls):
            return False

        @classmethod
        def can_use_node_types(cls):
            return False

        @classmethod
        def can_use_edge_types(cls):
            return False

        @classmethod
        def can_use_edge_type_features(cls):
            return False

        @classmethod
        def can_use_edge_features(cls):
            return False

    class RequiresEverything(Names, AbstractModel):  # the minimal valid deterministic model
        def __init__(self, random_state=None):
            super().__init__(random_state)

        @classmethod
        def is_stocastic(cls):
            return False

        @classmethod
        def requires_edge_weights(cls):
            return True

        @classmethod
        def requires_node_types(cls):
            return True

        @classmethod
        def requires_edge_types(cls):
            return True

        @classmethod
        def can_use_edge_type_features(cls):
            return False

        @classmethod
        def can_use_edge_features(cls):
            return False

    class UselessRequires(UsesNothing):  # says it cannot use weights AND answers whether it requires them
        @classmethod
        def requires_edge_weights(cls):
            return False

    class UselessCanUse(RequiresEverything):  # requires node types AND answers whether it can use them
        @classmethod
        def can_use_node_types(cls):
            return True

    class UselessIsUsing(UsesNothing):
        def is_using_edge_types(self):
            return False

    class UselessPositiveWeights(UsesNothing):
        @classmethod
        def requires_positive_edge_weights(cls):
            return False

    class SaysNothingAboutEdgeTypes(Names, AbstractModel):
        def __init__(self):
            super().__init__(7)

        @classmethod
        def is_stocastic(cls):
            return True

        @classmethod
        def can_use_edge_weights(cls):
            return False

        @classmethod
        def can_use_node_types(cls):
            return False

        @classmethod
        def can_use_edge_type_features(cls):
            return False

        @classmethod
        def can_use_edge_features(cls):
            return False

    class OptionalWeights(UsesNothing):  # can use weights, decides per instance
        @classmethod
        def can_use_edge_weights(cls):
            return True

        @classmethod
        def requires_edge_weights(cls):
            return False

        def is_using_edge_weights(self):
            return True

    return {
        "minimal stochastic model": (lambda: UsesNothing().parameters(), repr({"random_state": 42})),
        "stochastic without a seed": (lambda: UsesNothing(None), "ValueError"),
        "minimal deterministic model": (lambda: RequiresEverything().parameters(), repr({})),
        "deterministic with a seed": (lambda: RequiresEverything(3), "ValueError"),
        "requires_* beside can_use_* == False": (UselessRequires, "ValueError"),
        "can_use_* beside requires_* == True": (UselessCanUse, "ValueError"),
        "is_using_* beside can_use_* == False": (UselessIsUsing, "ValueError"),
        "requires_positive_edge_weights beside can_use_edge_weights == False": (UselessPositiveWeights, "ValueError"),
        "neither requires_ nor can_use_ for a capability": (SaysNothingAboutEdgeTypes, "ValueError"),
        "optional capability, all three implemented": (lambda: OptionalWeights().is_using_edge_weights(), "True"),
        "cannot use => does not require": (lambda: UsesNothing().requires_node_types(), "False"),
        "requires => can use": (lambda: RequiresEverything().can_use_edge_types(), "True"),
        "requires => is using": (lambda: RequiresEverything().is_using_edge_weights(), "True"),
        "requires => requires_positive_edge_weights is the subclass's call": (
            lambda: RequiresEverything().requires_positive_edge_weights(), "NotImplementedError"),
        "cannot use => weights need not be positive": (lambda: UsesNothing().requires_positive_edge_weights(), "False"),
        "is_using_* undecided by the class methods": (lambda: UsesNothing().is_using_node_types(), "NotImplementedError"),
        "task vocabulary is the subclass's": (lambda: UsesNothing().task_involves_topology(), "NotImplementedError"),
        "clone is the subclass's": (lambda: UsesNothing().clone(), "NotImplementedError"),
        "set_random_state on a deterministic model": (lambda: RequiresEverything().set_random_state(1), "ValueError"),
        "set_random_state": (lambda: (lambda m: (m.set_random_state(9), m.parameters())[1])(UsesNothing()),
                             repr({"random_state": 9})),
    }


def run_cases(AbstractModel):
    """name -> outcome actually observed."""
    observed = {}
    for name, (probe, _) in build_cases(AbstractModel).items():
        try:
            observed[name] = repr(probe())
        except Exception as error:  # the type is the outcome
            observed[name] = type(error).__name__
    return observed


def expected_outcomes():
    return {name: expected for name, (_, expected) in build_cases(object).items()}
