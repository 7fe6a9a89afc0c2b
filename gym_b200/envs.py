"""Host-side description of each env kind: spaces, constructor kwargs, reset options.

These mirror what the reference env classes set up in ``__init__`` (the spaces
an agent reads) and parse in ``reset(options=...)``; the dynamics themselves
live in ``csrc/envs.cuh``.
"""
import math
import warnings

import numpy as np

from gym_b200 import _lib
from gym_b200.spaces import Box, Discrete

_F32_MAX = np.finfo(np.float32).max


def _cartpole_spaces(params):
    # cartpole.py:100-117
    theta_threshold = 12 * 2 * math.pi / 360
    high = np.array([2.4 * 2, _F32_MAX, theta_threshold * 2, _F32_MAX], dtype=np.float32)
    return Box(-high, high, dtype=np.float32), Discrete(2)


def _mountaincar_spaces(params):
    # mountain_car.py:113-125
    low = np.array([-1.2, -0.07], dtype=np.float32)
    high = np.array([0.6, 0.07], dtype=np.float32)
    return Box(low, high, dtype=np.float32), Discrete(3)


def _mountaincar_cont_spaces(params):
    # continuous_mountain_car.py:121-140
    low = np.array([-1.2, -0.07], dtype=np.float32)
    high = np.array([0.6, 0.07], dtype=np.float32)
    return Box(low=low, high=high, dtype=np.float32), Box(low=-1.0, high=1.0, shape=(1,), dtype=np.float32)


def _pendulum_spaces(params):
    # pendulum.py:109-117
    high = np.array([1.0, 1.0, 8.0], dtype=np.float32)
    return Box(low=-high, high=high, dtype=np.float32), Box(low=-2.0, high=2.0, shape=(1,), dtype=np.float32)


def _acrobot_spaces(params):
    # acrobot.py:172-179
    high = np.array([1.0, 1.0, 1.0, 1.0, 4 * math.pi, 9 * math.pi], dtype=np.float32)
    return Box(low=-high, high=high, dtype=np.float32), Discrete(3)


def _lunar_spaces(params):
    # lunar_lander.py:248-292
    low = np.array([-1.5, -1.5, -5.0, -5.0, -math.pi, -5.0, -0.0, -0.0]).astype(np.float32)
    high = np.array([1.5, 1.5, 5.0, 5.0, math.pi, 5.0, 1.0, 1.0]).astype(np.float32)
    return Box(low, high), Discrete(4)


def _lunar_cont_spaces(params):
    # lunar_lander.py:285-287: Box(-1, +1, (2,), dtype=np.float32)
    return _lunar_spaces(params)[0], Box(-1, +1, (2,), dtype=np.float32)


def _walker_spaces(params):
    # bipedal_walker.py:192-237
    low = np.array([-math.pi, -5.0, -5.0, -5.0, -math.pi, -5.0, -math.pi, -5.0, -0.0, -math.pi, -5.0, -math.pi, -5.0,
                    -0.0] + [-1.0] * 10).astype(np.float32)
    high = np.array([math.pi, 5.0, 5.0, 5.0, math.pi, 5.0, math.pi, 5.0, 5.0, math.pi, 5.0, math.pi, 5.0, 5.0]
                    + [1.0] * 10).astype(np.float32)
    act = Box(np.array([-1, -1, -1, -1]).astype(np.float32), np.array([1, 1, 1, 1]).astype(np.float32))
    return Box(low, high), act


class KindInfo:
    def __init__(self, name, spaces, kwargs, bounds_keys, default_bounds, attrs, metadata):
        self.name = name
        self.spaces = spaces
        self.kwargs = kwargs            # accepted ctor kwargs -> (param slot, default)
        self.bounds_keys = bounds_keys  # reset(options=) keys, in the order of the C ABI's bounds[2]
        self.default_bounds = default_bounds
        self.attrs = attrs              # read-only attributes served by get_attr()/call()
        self.metadata = metadata


KINDS = {
    _lib.KIND_CARTPOLE: KindInfo(
        "CartPole", _cartpole_spaces, {}, ("low", "high"), (-0.05, 0.05),
        dict(gravity=9.8, masscart=1.0, masspole=0.1, total_mass=1.1, length=0.5, polemass_length=0.05,
             force_mag=10.0, tau=0.02, kinematics_integrator="euler",
             theta_threshold_radians=12 * 2 * math.pi / 360, x_threshold=2.4),
        {"render_modes": [], "render_fps": 50}),
    _lib.KIND_MOUNTAINCAR: KindInfo(
        "MountainCar", _mountaincar_spaces, {"goal_velocity": (0, 0)}, ("low", "high"), (-0.6, -0.4),
        dict(min_position=-1.2, max_position=0.6, max_speed=0.07, goal_position=0.5, force=0.001, gravity=0.0025),
        {"render_modes": [], "render_fps": 30}),
    _lib.KIND_MOUNTAINCAR_CONT: KindInfo(
        "MountainCarContinuous", _mountaincar_cont_spaces, {"goal_velocity": (0, 0)}, ("low", "high"), (-0.6, -0.4),
        dict(min_action=-1.0, max_action=1.0, min_position=-1.2, max_position=0.6, max_speed=0.07,
             goal_position=0.45, power=0.0015),
        {"render_modes": [], "render_fps": 30}),
    _lib.KIND_PENDULUM: KindInfo(
        "Pendulum", _pendulum_spaces, {"g": (0, 10.0)}, ("x_init", "y_init"), (math.pi, 1.0),
        dict(max_speed=8, max_torque=2.0, dt=0.05, m=1.0, l=1.0),
        {"render_modes": [], "render_fps": 30}),
    _lib.KIND_ACROBOT: KindInfo(
        "Acrobot", _acrobot_spaces, {}, ("low", "high"), (-0.1, 0.1),
        dict(dt=0.2, LINK_LENGTH_1=1.0, LINK_LENGTH_2=1.0, LINK_MASS_1=1.0, LINK_MASS_2=1.0,
             LINK_COM_POS_1=0.5, LINK_COM_POS_2=0.5, LINK_MOI=1.0, MAX_VEL_1=4 * math.pi, MAX_VEL_2=9 * math.pi,
             AVAIL_TORQUE=[-1.0, 0.0, +1], torque_noise_max=0.0, book_or_nips="book"),
        {"render_modes": [], "render_fps": 15}),
    _lib.KIND_LUNARLANDER: KindInfo(
        "LunarLander", _lunar_spaces, {}, ("low", "high"), (0.0, 0.0),
        dict(continuous=False, gravity=-10.0, enable_wind=False, wind_power=15.0, turbulence_power=1.5),
        {"render_modes": [], "render_fps": 50}),
    _lib.KIND_LUNARLANDER_CONT: KindInfo(
        "LunarLander", _lunar_cont_spaces, {}, ("low", "high"), (0.0, 0.0),
        dict(continuous=True, gravity=-10.0, enable_wind=False, wind_power=15.0, turbulence_power=1.5),
        {"render_modes": [], "render_fps": 50}),
    _lib.KIND_BIPEDALWALKER: KindInfo(
        "BipedalWalker", _walker_spaces, {}, ("low", "high"), (0.0, 0.0), dict(hardcore=False),
        {"render_modes": [], "render_fps": 50}),
    _lib.KIND_BIPEDALWALKER_HARDCORE: KindInfo(
        "BipedalWalker", _walker_spaces, {}, ("low", "high"), (0.0, 0.0), dict(hardcore=True),
        {"render_modes": [], "render_fps": 50}),
}


def verify_number_and_cast(x):
    """classic_control/utils.py:8-14"""
    try:
        return float(x)
    except (ValueError, TypeError):
        raise ValueError(f"An option ({x}) could not be converted to a float.")


def parse_reset_bounds(kind, options):
    """reset(options=) -> None (defaults) or the two doubles of the C ABI's `bounds`.

    low/high kinds follow maybe_parse_reset_bounds (classic_control/utils.py:17-46);
    Pendulum follows pendulum.py:143-152 (x_init / y_init, symmetric limits).
    """
    if options is None or kind in (_lib.KIND_LUNARLANDER, _lib.KIND_LUNARLANDER_CONT, _lib.KIND_BIPEDALWALKER,
                                   _lib.KIND_BIPEDALWALKER_HARDCORE):
        # their reset() ignores options
        return None
    info = KINDS[kind]
    k0, k1 = info.bounds_keys
    d0, d1 = info.default_bounds
    b0 = verify_number_and_cast(options.get(k0) if k0 in options else d0)
    b1 = verify_number_and_cast(options.get(k1) if k1 in options else d1)
    if info.bounds_keys == ("low", "high") and b0 > b1:
        raise ValueError(f"Lower bound ({b0}) must be lower than higher bound ({b1}).")
    return (b0, b1)


def resolve_variant(kind, kwargs):
    """ctor kwargs (gym.make(id, **kwargs)) -> (kind, flags) of b200gym_config: `continuous=True` selects the
    continuous-action kind (the reference registers it as LunarLanderContinuous-v2, gym/envs/__init__.py:62-68),
    `enable_wind=True` sets the wind flag."""
    flags = 0
    if kind in (_lib.KIND_LUNARLANDER, _lib.KIND_LUNARLANDER_CONT):
        if "continuous" in kwargs:
            kind = _lib.KIND_LUNARLANDER_CONT if kwargs["continuous"] else _lib.KIND_LUNARLANDER
        if kwargs.get("enable_wind", False):
            flags |= _lib.LUNAR_ENABLE_WIND
    if kind in (_lib.KIND_BIPEDALWALKER, _lib.KIND_BIPEDALWALKER_HARDCORE) and "hardcore" in kwargs:
        # gym/envs/__init__.py:79-85 registers hardcore=True as BipedalWalkerHardcore-v3
        kind = _lib.KIND_BIPEDALWALKER_HARDCORE if kwargs["hardcore"] else _lib.KIND_BIPEDALWALKER
    return kind, flags


def resolve_params(kind, kwargs):
    """ctor kwargs (gym.make(id, **kwargs)) -> the four doubles of b200gym_config.param."""
    info = KINDS[kind]
    params = [0.0, 0.0, 0.0, 0.0]
    if kind in (_lib.KIND_LUNARLANDER, _lib.KIND_LUNARLANDER_CONT):
        # LunarLander.__init__, lunar_lander.py:191-233
        values = dict(info.attrs)
        for key, value in kwargs.items():
            if key == "render_mode" and value is None:
                continue
            if key in ("wind_idx", "torque_idx"):   # engine extension: see B200VectorEnv
                continue
            if key not in info.attrs:
                raise TypeError(f"LunarLander got an unexpected keyword argument '{key}'")
            values[key] = value
        gravity = float(values["gravity"])
        assert -12.0 < gravity and gravity < 0.0, f"gravity (current value: {gravity}) must be between -12 and 0"
        wind_power, turbulence_power = float(values["wind_power"]), float(values["turbulence_power"])
        if 0.0 > wind_power or wind_power > 20.0:
            warnings.warn(f"WARN: wind_power value is recommended to be between 0.0 and 20.0, (current value: {wind_power})")
        if 0.0 > turbulence_power or turbulence_power > 2.0:
            warnings.warn("WARN: turbulence_power value is recommended to be between 0.0 and 2.0, "
                          f"(current value: {turbulence_power})")
        return [gravity, wind_power, turbulence_power, 0.0]
    if kind in (_lib.KIND_BIPEDALWALKER, _lib.KIND_BIPEDALWALKER_HARDCORE):
        for key, value in kwargs.items():   # BipedalWalker.__init__(render_mode, hardcore), bipedal_walker.py:167
            if key == "render_mode" and value is None:
                continue
            if key != "hardcore":
                raise TypeError(f"BipedalWalker got an unexpected keyword argument '{key}'")
        return params
    for key, (slot, default) in info.kwargs.items():
        params[slot] = float(default)
    for key, value in kwargs.items():
        if key == "render_mode":
            if value is not None:
                raise ValueError("gym_b200 does not render (render_mode must be None)")
            continue
        if key not in info.kwargs:
            raise TypeError(f"{info.name} got an unexpected keyword argument '{key}'")
        params[info.kwargs[key][0]] = float(value)
    return params
