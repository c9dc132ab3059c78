import sys

from .host import main

sys.exit(main())
