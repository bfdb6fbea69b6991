class Data:          # type-hint only in memory/ltm.py and memory/replay_buffer.py of the reference
    pass
