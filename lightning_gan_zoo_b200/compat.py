"""Expose the B200 modules under the reference's import paths.

The reference instantiates its networks from Hydra `_target_` strings
(`core.models.hologan_generator.Generator`, `core.models.hologan_discriminator.Discriminator`,
conf/expt/hologan.yaml:25-31).  `install()` registers this package's mirrors under those dotted
names -- the two networks, `core.utils.hologan.create_hologan_lr_scheduler` (conf/lr_scheduler/hologan.yaml) and
`core.lightning_module.HOLOGAN` (model.lm) -- so `+expt=hologan` resolves to the B200 path without editing the YAML.
"""
from __future__ import annotations

import importlib
import sys
import types

_ALIASES = {
    "core.models.hologan_generator": "lightning_gan_zoo_b200.core.models.hologan_generator",
    "core.models.hologan_discriminator": "lightning_gan_zoo_b200.core.models.hologan_discriminator",
    "core.utils.hologan": "lightning_gan_zoo_b200.core.utils.hologan",          # conf/lr_scheduler/hologan.yaml
    "core.lightning_module": "lightning_gan_zoo_b200.core.lightning_module",    # model.lm of conf/expt/hologan.yaml
}


def install(force: bool = False) -> None:
    for pkg in ("core", "core.models", "core.utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []          # mark as package
            sys.modules[pkg] = m
    for alias, target in _ALIASES.items():
        if alias in sys.modules and not force:
            continue
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        parent, _, leaf = alias.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    setattr(sys.modules["core"], "models", sys.modules["core.models"])
    setattr(sys.modules["core"], "utils", sys.modules["core.utils"])
