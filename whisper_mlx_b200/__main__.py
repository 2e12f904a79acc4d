from whisper_mlx_b200.cli import main

main()
