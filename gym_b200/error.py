"""Exception names the engine shares with the reference (gym/error.py:5-193).

Only the classes that can surface on the step()/reset() path are restated; the
names match so `except gym.error.ResetNeeded` style code keeps its shape.
"""


class Error(Exception):
    """Base class (gym/error.py:5)."""


class UnregisteredEnv(Error):
    """Unknown environment id (gym/error.py:13)."""


class NameNotFound(UnregisteredEnv):
    """gym/error.py:21."""


class VersionNotFound(UnregisteredEnv):
    """gym/error.py:25."""


class ResetNeeded(Error):
    """step() before reset() (gym/error.py:45, raised at gym/wrappers/order_enforcing.py:33-37)."""


class InvalidAction(Error):
    """Action outside the action space (gym/error.py:57)."""


class DependencyNotInstalled(Error):
    """The CUDA extension is missing / cannot run: the engine has no CPU fallback."""


class AlreadyPendingCallError(Error):
    """reset_async/step_async while another call is pending (gym/error.py:143)."""

    def __init__(self, message, name):
        super().__init__(message)
        self.name = name


class NoAsyncCallError(Error):
    """step_wait/reset_wait without a pending call (gym/error.py:153)."""

    def __init__(self, message, name):
        super().__init__(message)
        self.name = name


class ClosedEnvironmentError(Error):
    """Operation on a closed vector env (gym/error.py:163, raised at gym/vector/async_vector_env.py:518-522)."""
