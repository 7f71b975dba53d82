"""fp16 conversion helpers of the sampling path (guided_diffusion/fp16_util.py:15-32).  The reference walks
nn.Conv modules; here a model knows which of its parameters are torso convolutions, so the two functions
delegate to it.  The master-parameter / loss-scaling trainer (fp16_util.py:35-236) is training-only."""


def convert_module_to_f16(model):
    model.convert_to_fp16()


def convert_module_to_f32(model):
    model.convert_to_fp32()
