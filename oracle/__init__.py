"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference forward path and the harness that
imports the real reference to pin it.  Nothing under the product package may import this."""
