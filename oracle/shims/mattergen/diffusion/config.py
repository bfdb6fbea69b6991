class Config:
    pass
