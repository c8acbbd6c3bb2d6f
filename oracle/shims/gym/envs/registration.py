from gym import register  # noqa: F401  (reference: crowd_sim/__init__.py:1)
