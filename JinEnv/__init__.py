"""Drop-in package: ``from JinEnv import JinEnv`` like the reference."""
