import json, sys
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)['preprocess']
        print("gray %.1f us frac %.3f | bgr %.1f us frac %.3f | bad %d" % (d['gray']['ms'] * 1e3, d['gray']['roofline']['frac'], d['bgr']['ms'] * 1e3, d['bgr']['roofline']['frac'], d['bad_boxes']))
