/* empty stand-in for <GL/glu.h> (headless harness) */
