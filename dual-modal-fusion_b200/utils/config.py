"""config.yml loader with the reference's entry point get_render_config(path)
(utils/config.py:12-41): YAML -> Jinja2 render with {{dc}},{{tr}},{{ep}},{{bs}},{{mn}},{{FN}},{{num}},
{{p2}} -> YAML, Categories_Number = len(color table) (background included), result directories.

Differences, all additive: the template is read from `path` (the reference re-opens 'config.yml' in
the CWD), the dqtl block is optional (the shipped config.yml has none and the reference loader
raises KeyError on it), and stale result directories are only deleted when `delete: 1`."""
import os
import shutil
from pathlib import Path

import yaml
from jinja2 import Template


def get_config(path):
    with open(path, encoding='utf-8') as f:
        return yaml.load(f, yaml.FullLoader)


def get_render_config(path):
    data = get_config(path)
    dqtl = data.get('dqtl') or {}
    with open(path, 'r', encoding='utf-8') as f:
        template = Template(f.read())
    rendered = template.render(
        parameter1='value1', p2=Path(__file__).resolve().parent.parent, dc=data['data_city'],
        num=len(data['DATA_DICT'][data['data_city']]['color']), tr=data['train_rate'], ep=data['epoch'],
        bs=data['batchsize'], expo_result=data['expo_result'], parameters=data['parameters'], mn=data['model_name'],
        FN=data['FILE_NUM'], ne=dqtl.get('num_epochs', 0), ps=dqtl.get('pic_size', 0))
    return get_dump_config(yaml.load(rendered, yaml.FullLoader))


def get_dump_config(y):
    os.makedirs(y['RESULT'], exist_ok=True)
    stem = y['RESULT'] + y['model_name'] + "__"
    filenum = 0
    if y['train']['index'] != 0:
        while os.path.exists(stem + str(filenum) + '_result.xlsx') or os.path.exists(stem + str(filenum) + '_output/'):
            filenum += 1
        if y.get('delete'):
            for num in range(filenum - 1, -1, -1):      # unfinished runs (output dir without a result sheet)
                out_dir = stem + str(num) + '_output'
                if os.path.isdir(out_dir) and not os.path.isfile(stem + str(num) + '_result.xlsx'):
                    shutil.rmtree(out_dir)
                    filenum = num
        y['FILE_NUM'] = filenum
    else:
        filenum = y['FILE_NUM']
    y['RESULT_excel'] = stem + str(filenum) + '_result.xlsx'
    y['RESULT_output'] = stem + str(filenum) + '_output/'
    y['schedule']['lr'] = float(y['schedule']['lr'])
    y['schedule']['base_lr'] = float(y['schedule']['base_lr'])
    y['Categories_Number'] = int(y['Categories_Number'])
    if isinstance(y.get('dqtl'), dict):
        for key in ('lr', 'tao', 'epsilon'):
            if key in y['dqtl']:
                y['dqtl'][key] = float(y['dqtl'][key])
    y = yaml.safe_load(yaml.dump(y))
    if y['train']['save_best']:
        os.makedirs(y['RESULT_output'], exist_ok=True)
    return y


if __name__ == "__main__":
    print(get_render_config("./config.yml"))
