class Config(dict):  # reference: torch2scripts.py:5 (imported, unused on the paths exercised here)
    pass
