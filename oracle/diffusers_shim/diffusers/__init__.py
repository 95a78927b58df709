"""Two-symbol stand-in for ``diffusers`` so that the reference's ``attention.py``
(`from diffusers.models.lora import LoRALinearLayer`, `attention.py:4`;
`from diffusers.utils.import_utils import is_xformers_available`, `attention.py:6`)
imports verbatim in a container without diffusers.  Test infrastructure only."""
