"""Same command line as the reference's DiffPhar/generate_phars.py; see cmd_gen_b200/generate_phars.py."""
from cmd_gen_b200.generate_phars import main

if __name__ == "__main__":
    main()
