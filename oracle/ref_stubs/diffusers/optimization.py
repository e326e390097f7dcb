def get_cosine_schedule_with_warmup(*args, **kwargs):
    raise NotImplementedError("stub: diffusers is not installed; training is out of scope")
