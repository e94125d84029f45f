from .command_line_interface import main

main()
