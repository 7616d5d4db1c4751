"""Error types mirroring tgm/exceptions.py (same names, same meaning) so user code that
catches the reference's exceptions keeps working."""


class TGMError(Exception):
    """Root of all library errors."""


class BadHookProtocolError(TGMError):
    """A hook object does not satisfy the DGHook protocol."""


class BadEncoderProtocolError(TGMError):
    """An encoder passed to HookManager.validate_requirement lacks `requires`/__call__."""


class UnresolvableHookDependenciesError(TGMError):
    """No execution order satisfies the hooks' requires/produces sets."""


class InvalidNodeIDError(TGMError):
    """A node id collides with PADDED_NODE_ID or lies outside the graph's id range."""


class EmptyGraphError(TGMError):
    """A graph without any event was requested."""


class EventOrderedConversionError(TGMError):
    """A time-unit operation was attempted on an event-ordered ('r') graph."""


class InvalidDiscretizationError(TGMError):
    """Batch unit finer than the graph's own time granularity."""


class EmptyBatchError(TGMError):
    """An empty batch was met while on_empty='raise'."""
