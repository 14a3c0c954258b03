"""Drop-in package: ``from PDP import PDP`` like the reference (Examples/IRL/quadrotor/uav_PDP.py:1)."""
