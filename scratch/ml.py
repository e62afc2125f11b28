import sys, json; sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import measure_configs as m
print(json.dumps(m.multilayer_config(), indent=1))
