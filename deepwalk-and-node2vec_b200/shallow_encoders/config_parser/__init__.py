from shallow_encoders.config_parser.core import GlobalConfig, load_config  # noqa: F401
