"""dm_control_b200 — B200-native batched forward-dynamics step behind dm_control's Physics.step()."""
__version__ = '0.1.0'
