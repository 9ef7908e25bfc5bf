"""Single-process stand-in for the collective backend: the degenerate world of ONE rank.  It exposes exactly the names
the trainer, the callbacks and the data loader look up on `hvd` (comm convention) and answers them the
way `polus_b200.comm` does when WORLD_SIZE == 1, so code written against either backend runs unchanged."""


class _OneRankWorld:
    """A world of size 1: every collective is the identity."""
    WORLD = 1

    def init(self):
        return "mock"

    def rank(self):
        return self.WORLD - 1

    local_rank = rank

    def size(self):
        return self.WORLD

    def DistributedGradientTape(self, tape, **unused):
        # averaging gradients over one rank leaves them as they are: hand the tape back
        return tape

    def broadcast_variables(self, variables, root_rank=0):
        assert root_rank == 0, "a world of one rank has no other root"

    def allgather_object(self, obj):
        return [obj] * self.WORLD


_world = _OneRankWorld()
init, rank, local_rank, size = _world.init, _world.rank, _world.local_rank, _world.size
DistributedGradientTape = _world.DistributedGradientTape
broadcast_variables = _world.broadcast_variables
allgather_object = _world.allgather_object
